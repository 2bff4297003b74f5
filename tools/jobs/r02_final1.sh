#!/bin/bash
# Round-2 final one-GPU evidence: tests, bench lines (default / exact), reference arm, launch list, ncu summaries, K4, sanitizers
mkdir -p gpurun_out
O=gpurun_out
T=${1:-f1}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python bench.py > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
python bench.py --exact 2 --no-cpu > $O/r02_${T}_bench_exact2.json 2> $O/r02_${T}_bench_exact2.err
python bench.py --exact 1 --no-cpu > $O/r02_${T}_bench_exact1.json 2> $O/r02_${T}_bench_exact1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_${T}_ref.json 2> $O/r02_${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k3|k_ec_lik|k_ec_comb' -c 60 --csv --log-file $O/r02_${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r02_${T}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k3|k_ec_lik|k_ec_comb' -s 8 -c 8 -o $O/r02_${T}_step python tools/ec_probe.py --reps 2 --steps 3 > $O/r02_${T}_step_ncu.log 2>&1
python tools/bench_hsb.py --reps 5 > $O/r02_${T}_hsb.jsonl 2> $O/r02_${T}_hsb.err
timeout 300 ncu --set full --clock-control none -k regex:'k4_' -c 12 -o $O/r02_${T}_k4 python tools/bench_hsb.py --reps 1 --cpu-threads 0 > $O/r02_${T}_k4_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_${T}_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_${T}_racecheck.log 2>&1
tail -3 $O/r02_${T}_pytest.log; cut -c1-250 $O/r02_${T}_bench.json; cut -c1-200 $O/r02_${T}_bench_exact2.json
