"""world_size-2 gloo test (CPU) of the N>1 host logic: equal-nnz row partitioning + one all-reduce(sum) of the
transcript-length gradient per step reproduces the single-rank gradient."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import Fixture
    from oracle import polee_oracle as O
    import polee_b200 as pb
    fx = Fixture()
    sample = pb.RNASeqSample(fx.m, fx.n, fx.colptr, fx.rowval, fx.nzval, fx.efflens)
    bounds = pb.partition_rows(sample, world)
    blk = pb.api.row_block(sample, int(bounds[rank]), int(bounds[rank + 1]))
    xs = np.random.default_rng(5).dirichlet(np.ones(fx.n)).astype(np.float32)
    lp, g = O.Model(blk.m, blk.n, blk.colptr, blk.rowval, blk.nzval).log_likelihood(xs, gradonly=False)
    buf = torch.from_numpy(np.concatenate([g, [lp]]))
    dist.all_reduce(buf)                       # the one collective of a step
    if rank == 0:
        lp_full, g_full = O.Model(fx.m, fx.n, fx.colptr, fx.rowval, fx.nzval).log_likelihood(xs, gradonly=False)
        out["g_err"] = float(np.max(np.abs(buf[:-1].numpy() - g_full) / np.abs(g_full).clip(1e-300)))
        out["lp_err"] = abs(float(buf[-1]) - lp_full) / abs(lp_full)
        out["nnz"] = [int(len(blk.rowval))]
    dist.barrier()
    dist.destroy_process_group()


def test_row_partitioned_gradient_allreduce_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["g_err"] < 1e-12 and out["lp_err"] < 1e-13
