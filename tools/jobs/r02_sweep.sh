#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_k3_sweep.log
: > $O
for B in 256 512 1024; do for V in 0 1 2 3; do
  echo "== bin $B variant $V" >> $O
  POLEE_TREE_BIN_NODES=$B POLEE_TREE_VARIANT=$V timeout 120 python tools/ec_probe.py --reps 30 2>&1 | grep "likelihood pass" >> $O
done; done
echo "== paths fwd" >> $O; POLEE_TREE_FWD=paths timeout 120 python tools/ec_probe.py --reps 30 2>&1 | grep "likelihood pass" >> $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $O
cat $O
