#!/bin/bash
# gpu tests (all, no -x), bench line, K4 device-resident bench, launch list
mkdir -p gpurun_out
O=gpurun_out
T=${1:-j3}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python bench.py > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
python tools/bench_hsb.py --reps 5 > $O/r02_${T}_hsb.jsonl 2> $O/r02_${T}_hsb.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_${T}_launches.log 2>&1
tail -15 $O/r02_${T}_pytest.log; cut -c1-400 $O/r02_${T}_bench.json; cat $O/r02_${T}_hsb.jsonl | cut -c1-200
