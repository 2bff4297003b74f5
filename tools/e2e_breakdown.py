"""Where the 0.66 s of one approximate_likelihood call (C3, host buffers) goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polee_b200 as pb
from polee_b200 import synth
from bench import generate
s, tree, K = generate("c3", "cuda:0")
m, n = s["m"], s["n"]
pin = lambda t, dt: torch.empty(t.shape, dtype=dt, pin_memory=True).copy_(t.to(dt)).numpy()
colptr = pin(s["colptr"], torch.int32).view(np.uint32); rowval = pin(s["rowval"], torch.int32).view(np.uint32); nzval = pin(s["nzval"], torch.float32)
eff = s["efflens"].cpu().numpy(); del s; torch.cuda.empty_cache()
sample = pb.RNASeqSample(m, n, colptr, rowval, nzval, eff)
for rep in range(3):
    T = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = pb.Handle(num_mc_samples=K, num_steps=500); T["create"] = time.perf_counter() - t0; t = time.perf_counter()
    h.check(h.lib.polee_set_matrix_csc(h.h, m, n, colptr.ctypes.data_as(pb.api._P), rowval.ctypes.data_as(pb.api._P), nzval.ctypes.data_as(pb.api._P), None)); h.m, h.n = m, n
    T["set_matrix (H2D + layout build)"] = time.perf_counter() - t; t = time.perf_counter()
    h.set_efflens(eff); T["set_efflens"] = time.perf_counter() - t; t = time.perf_counter()
    h.set_tree(*tree); T["set_tree"] = time.perf_counter() - t; t = time.perf_counter()
    h.init_params(); h.run_steps(500); T["enqueue 500 steps"] = time.perf_counter() - t; t = time.perf_counter()
    h.sync(); T["wait"] = time.perf_counter() - t; t = time.perf_counter()
    h.get_params(); T["get_params"] = time.perf_counter() - t; t = time.perf_counter()
    h.close(); T["destroy"] = time.perf_counter() - t
    print(rep, "total %.3f" % (time.perf_counter() - t0), {k: round(v * 1e3, 1) for k, v in T.items()})
