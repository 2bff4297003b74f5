#!/bin/bash
# One-GPU job: gpu tests, cold-fit breakdown, K4 device-resident bench + ncu, ncu of the K3 kernels, racecheck
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_j2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_j2_pytest.log
POLEE_SETUP_TIMING=1 python tools/e2e_breakdown.py > $O/r02_j2_e2e_breakdown.log 2>&1
python tools/bench_hsb.py --reps 5 > $O/r02_j2_hsb.jsonl 2> $O/r02_j2_hsb.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k4_' -c 14 -o $O/r02_j2_k4 python tools/bench_hsb.py --reps 1 --cpu-threads 0 > $O/r02_j2_k4_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3' -s 6 -c 12 -o $O/r02_j2_k3 python tools/ec_probe.py --reps 2 --steps 3 > $O/r02_j2_k3_ncu.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_j2_racecheck.log 2>&1
tail -3 $O/r02_j2_pytest.log; cat $O/r02_j2_hsb.jsonl; tail -8 $O/r02_j2_e2e_breakdown.log
