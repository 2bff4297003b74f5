#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 240 $RUN tools/check_multi_gpu.py > $O/r02_n2b_check.log 2>&1; echo "rc=$?" >> $O/r02_n2b_check.log
timeout 400 $RUN bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_n2b_bench.json 2> $O/r02_n2b_bench.err; echo "rc=$?" >> $O/r02_n2b_bench.err
tail -3 $O/r02_n2b_check.log; grep '^{' $O/r02_n2b_bench.json | cut -c1-200
