#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T=${1:-p1}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
tail -15 $O/r02_${T}_pytest.log | cut -c1-300
