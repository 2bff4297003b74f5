#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T=${1:-j6}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python bench.py --no-cpu > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k3|k_ec_lik|k_ec_comb' -c 120 --csv --log-file $O/r02_${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_${T}_launches.log 2>&1
tail -6 $O/r02_${T}_pytest.log | cut -c1-200; cut -c1-300 $O/r02_${T}_bench.json
