#!/bin/bash
# final validation of the round: gpu tests, smoke, the default bench line, the warm set-up phases after the create fix
mkdir -p gpurun_out
O=gpurun_out
T=${1:-f2}
python bench.py > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_${T}_smoke.log
POLEE_SETUP_TIMING=1 python tools/e2e_breakdown.py --reps 3 > $O/r02_${T}_e2e_alt.log 2>&1
tail -3 $O/r02_${T}_pytest.log | cut -c1-200; tail -2 $O/r02_${T}_smoke.log; grep '^{' $O/r02_${T}_bench.json | cut -c1-200; grep total $O/r02_${T}_e2e_alt.log | cut -c1-300
