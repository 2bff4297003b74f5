#!/bin/bash
# bench line at N GPUs of one box (N = 1: plain python), as the driver launches it
N=${1:-1}
mkdir -p gpurun_out
O=gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_scale_${N}.json 2> $O/r02_scale_${N}.err
  python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_scale_${N}.json 2> $O/r02_scale_${N}.err
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/r02_scale_${N}_ref.json 2> $O/r02_scale_${N}_ref.err
fi
grep '^{' $O/r02_scale_${N}.json | cut -c1-200; tail -2 $O/r02_scale_${N}.err
