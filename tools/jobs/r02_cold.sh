#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_cold.log
: > $O
for i in 1 2 3; do
  echo "== child alone $i" >> $O
  POLEE_SETUP_TIMING=1 python bench.py --oneshot-child 2>&1 | grep -E "oneshot|alloc|set_matrix|inputs resident|schedule|initial mu" >> $O
done
echo "== child while a parent process holds a context and has just freed 8 GB" >> $O
python - >> $O 2>&1 <<'PY'
import torch, subprocess, sys, os
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda"); x.fill_(1); torch.cuda.synchronize()
del x; torch.cuda.empty_cache()
env = dict(os.environ, POLEE_SETUP_TIMING="1")
for i in range(2):
    out = subprocess.run([sys.executable, "bench.py", "--oneshot-child"], env=env, capture_output=True, text=True)
    print([l for l in (out.stdout + out.stderr).splitlines() if "oneshot" in l or "alloc" in l])
PY
cat $O
