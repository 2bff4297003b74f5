#!/bin/bash
# the bench as the FIRST command on a fresh box (what the driver does), with the set-up phase marks of its one-shot children
mkdir -p gpurun_out
O=gpurun_out
T=${1:-f3}
POLEE_SETUP_TIMING=1 python bench.py --no-cpu > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err
grep '^{' $O/r02_${T}_bench.json | cut -c1-200; grep -v "setup\] ec" $O/r02_${T}_bench.err | head -60
