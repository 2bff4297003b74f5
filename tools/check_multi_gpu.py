"""torchrun --nproc-per-node N tools/check_multi_gpu.py: the row-partitioned fit (one all-reduce per step inside the
captured graph) against the single-GPU fit of the whole sample -- same Philox noise on every rank, so the parameters
must agree up to the summation order of the gradient: <= 1e-5 after 3 ADAM steps (ADAM normalises the gradient, so
rounding differences grow with the step count: ~1e-3 after 60 steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import polee_b200 as pb
from polee_b200 import synth
from polee_b200 import api as pbapi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = synth.make_sample(400000, 20000, seed=77)
ns = synth.to_numpy_sample(s)
tree = synth.balanced_tree(20000, s["gene_sizes"].numpy())
whole = pb.RNASeqSample(ns["m"], ns["n"], ns["colptr"], ns["rowval"], ns["nzval"], ns["efflens"])
steps = int(os.environ.get("CHECK_STEPS", "3"))
bounds = pb.partition_rows(whole, world)
block = pbapi.row_block(whole, bounds[rank], bounds[rank + 1])
h = pb.Handle(device=local, num_steps=steps, num_mc_samples=8, seed=4242)
h.set_sample(block)
h.set_tree(*tree)
allreduce = pbapi.connect_ranks(h, dist)
h.init_params()
h.run_steps(steps)
h.sync()
mu, om, al = h.get_params()
h.close()
ref = pb.approximate_likelihood(pb.LogitSkewNormalPTTApprox(), whole, tree_topology=tree, num_steps=steps,
                                num_mc_samples=8, seed=4242, device=local)
d = max(np.abs(mu - ref["mu"]).max(), np.abs(om - ref["omega"]).max(), np.abs(al - ref["alpha"]).max())
t = torch.tensor([d], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
g = [None] * world
dist.all_gather_object(g, float(np.abs(mu).sum()))
if rank == 0:
    print("ranks %d: max |param - single GPU| = %.3e ; identical across ranks: %s" % (world, t.item(), len(set(g)) == 1))
    assert t.item() <= (1e-5 if steps <= 3 else 5e-3) and len(set(g)) == 1
dist.barrier()
dist.destroy_process_group()
