#!/bin/bash
# validation of polee_set_sample (one-call set-up) + the K = 6 line + the opt-in DFS backward kernel's parity
mkdir -p gpurun_out
O=gpurun_out
T=${1:-e1}
python -m pytest tests -m gpu -q > $O/r02_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest.log
python bench.py --no-cpu > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
POLEE_SETUP_TIMING=1 python tools/e2e_breakdown.py --one-call > $O/r02_${T}_e2e_onecall.log 2>&1
POLEE_SETUP_TIMING=1 python tools/e2e_breakdown.py > $O/r02_${T}_e2e_threecalls.log 2>&1
python bench.py --config c3-k6 --no-cpu --no-e2e > $O/r02_${T}_bench_k6.json 2> $O/r02_${T}_bench_k6.err
POLEE_TREE_BWD=dfs python -m pytest tests -m gpu -q -k "tree_forward_backward or one_step_of_draws or fit_trajectory or hsb" > $O/r02_${T}_pytest_dfs_bwd.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_dfs_bwd.log
tail -4 $O/r02_${T}_pytest.log | cut -c1-300; tail -2 $O/r02_${T}_pytest_dfs_bwd.log; grep '^{' $O/r02_${T}_bench.json | cut -c1-120; grep total $O/r02_${T}_e2e_onecall.log $O/r02_${T}_e2e_threecalls.log | cut -c1-330; grep '^{' $O/r02_${T}_bench_k6.json | cut -c1-160
