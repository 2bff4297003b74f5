#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T=${1:-q1}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python bench.py --no-cpu --no-e2e > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
tail -3 $O/r02_${T}_pytest.log; grep '^{' $O/r02_${T}_bench.json | cut -c1-200
