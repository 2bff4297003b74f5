#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T=${1:-j5}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python tools/bench_hsb.py --reps 5 > $O/r02_${T}_hsb.jsonl 2> $O/r02_${T}_hsb.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3' -s 6 -c 6 -o $O/r02_${T}_k3 python tools/ec_probe.py --reps 2 --steps 3 > $O/r02_${T}_k3_ncu.log 2>&1
tail -4 $O/r02_${T}_pytest.log; cat $O/r02_${T}_hsb.jsonl | cut -c1-200
