#!/bin/bash
# N-GPU job: multi-rank correctness, per-step times with both all-reduces, the bench line
N=${1:-2}
T=${2:-n2}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ "$N" -le 2 ] && timeout 240 $RUN tools/check_multi_gpu.py > $O/r02_${T}_check.log 2>&1; echo "rc=$?" >> $O/r02_${T}_check.log
timeout 240 $RUN tools/step_times.py --steps 60 > $O/r02_${T}_steps_peer.log 2>&1; echo "rc=$?" >> $O/r02_${T}_steps_peer.log
POLEE_ALLREDUCE=nccl timeout 240 $RUN tools/step_times.py --steps 60 > $O/r02_${T}_steps_nccl.log 2>&1; echo "rc=$?" >> $O/r02_${T}_steps_nccl.log
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_${T}_bench.json 2> $O/r02_${T}_bench.err; echo "rc=$?" >> $O/r02_${T}_bench.err
tail -3 $O/r02_${T}_check.log; tail -2 $O/r02_${T}_steps_peer.log; tail -2 $O/r02_${T}_steps_nccl.log; tail -c 1500 $O/r02_${T}_bench.json; tail -3 $O/r02_${T}_bench.err
