#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_last_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_last_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_last_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_last_smoke.log
python bench.py > $O/r02_last_bench.json 2> $O/r02_last_bench.err
tail -3 $O/r02_last_pytest.log; tail -2 $O/r02_last_smoke.log; grep '^{' $O/r02_last_bench.json | cut -c1-200
