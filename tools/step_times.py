"""Per-step device times of the row-partitioned fit (run under torch.distributed.run): one CUDA event per ADAM step on
the handle's stream, rank 0 prints the series (max over ranks) -- shows warm-up effects the mean of a short window hides."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import torch.distributed as dist
import polee_b200 as pb
from polee_b200 import api as pbapi
from bench import generate, equal_nnz_bounds, row_block_device

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--steps", type=int, default=80)
ap.add_argument("--chunk", type=int, default=1)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = "cuda:%d" % local
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
s, tree, K = generate(a.config, dev)
b = equal_nnz_bounds(s, world)
m_loc, cp, rv, nz = row_block_device(s, b[rank], b[rank + 1])
efflens = s["efflens"].cpu().numpy()
n = s["n"]
del s
h = pb.Handle(device=local, num_mc_samples=K, num_steps=a.steps * a.chunk + 8)
h.set_matrix_device(m_loc, n, cp.data_ptr(), rv.data_ptr(), nz.data_ptr())
h.set_efflens(efflens)
h.set_tree(*tree)
if world > 1:
    allreduce = pbapi.connect_ranks(h, dist)
h.init_params()
stream = torch.cuda.ExternalStream(h.stream(), device=dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
ev[0].record(stream)
for i in range(a.steps):
    h.run_steps(a.chunk)
    ev[i + 1].record(stream)
h.sync()
torch.cuda.synchronize()
t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) / a.chunk for i in range(a.steps)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    v = t.cpu().numpy()
    print("ranks %d, per-step ms (max over ranks), chunk %d:" % (world, a.chunk))
    print(" ".join("%.3f" % x for x in v))
    print("median %.4f, mean first 20 %.4f, mean last 20 %.4f" % (np.median(v), v[:20].mean(), v[-20:].mean()))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
