#!/bin/bash
# One-GPU validation job: gpu tests, the default bench line, a launch list and a full ncu capture of one step.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_v_pytest.log
python bench.py > $O/r02_v_bench.json 2> $O/r02_v_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_v_ref.json 2> $O/r02_v_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_v_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02_v_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ec|k3|k_fused|k2_|k1_|k_gene|k_narrow|k_widen' -c 24 -o $O/r02_v_step python tools/ec_probe.py --steps 2 > $O/r02_v_step_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_v_memcheck.log 2>&1
tail -3 $O/r02_v_pytest.log; cat $O/r02_v_bench.json | cut -c1-600
