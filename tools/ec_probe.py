"""Likelihood-pass probe: build the device layouts of a config, print what they hold, time the pass (CUDA events,
polee_time_kernel) and K3, and compare one loglik_grad with the oracle on two draws (--check)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import polee_b200 as pb
from bench import generate

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--check", action="store_true")
ap.add_argument("--steps", type=int, default=0, help="also run this many uncaptured ADAM steps (for ncu)")
a = ap.parse_args()
s, tree, K = generate(a.config, "cuda:0")
m, n = s["m"], s["n"]
h = pb.Handle(num_mc_samples=K, num_steps=max(a.steps, 1), use_cuda_graph=False)
cp, rv, nz = s["colptr"].to(torch.int32), s["rowval"].to(torch.int32), s["nzval"].contiguous()
torch.cuda.synchronize()
t0 = time.perf_counter()
h.set_matrix_device(m, n, cp.data_ptr(), rv.data_ptr(), nz.data_ptr())
torch.cuda.synchronize()
print("set_matrix %.1f ms" % ((time.perf_counter() - t0) * 1e3))
h.set_efflens(s["efflens"].cpu().numpy())
h.set_tree(*tree)
info = h.layout_info()
print("layout", info, "nnz", s["nnz"])
st = h.step_stats()
print("step_stats", st)
h.time_kernel(1, 2)   # first launches load the kernels lazily
h.time_kernel(3, 2)
t1 = h.time_kernel(1, a.reps)
t3 = h.time_kernel(3, a.reps)
print("likelihood pass %.4f ms (%.1f GB/s moved), K3 %.4f ms" % (t1, st["bytes_k1"] / t1 / 1e6, t3))
if a.check:
    from polee_b200 import synth
    from oracle import polee_oracle as O
    ns = synth.to_numpy_sample(s)
    xs = np.random.default_rng(0).dirichlet(np.ones(n), K).astype(np.float32).clip(1e-10)
    lp, g = h.loglik_grad(xs, gradonly=False)
    M = O.Model(m, n, ns["colptr"], ns["rowval"], ns["nzval"])
    for k in sorted({0, K - 1}):
        t0 = time.perf_counter()
        lp_o, g_o = M.log_likelihood(xs[k], gradonly=False)
        dt = time.perf_counter() - t0
        nzm = g_o != 0
        print("draw %d: lp relerr %.3e, x_grad max relerr %.3e, zero cols equal %s (oracle %.2f s)" % (
            k, abs(lp[k] - lp_o) / abs(lp_o), float(np.max(np.abs(g[k][nzm] - g_o[nzm]) / g_o[nzm])),
            bool(np.array_equal(g[k][~nzm], g_o[~nzm])), dt))
if a.steps:
    h.init_params()
    h.run_steps(a.steps)
    h.sync()
print("done")
