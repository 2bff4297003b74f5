"""Where the time of one approximate_likelihood call (C3, host buffers) goes: the three set-up calls against
polee_set_sample (one call, tree host work on a second thread), alternating in one process, gc collected before each."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polee_b200 as pb
from bench import generate
REPS = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 4
s, tree, K = generate("c3", "cuda:0")
m, n = s["m"], s["n"]
pin = lambda t, dt: torch.empty(t.shape, dtype=dt, pin_memory=True).copy_(t.to(dt)).numpy()
colptr = pin(s["colptr"], torch.int32).view(np.uint32); rowval = pin(s["rowval"], torch.int32).view(np.uint32); nzval = pin(s["nzval"], torch.float32)
eff = s["efflens"].cpu().numpy(); del s; torch.cuda.empty_cache()
sample = pb.RNASeqSample(m, n, colptr, rowval, nzval, eff)
for rep in range(2 * REPS):
    one_call = rep % 2 == 1
    gc.collect()
    T = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = pb.Handle(num_mc_samples=K, num_steps=500); T["create"] = time.perf_counter() - t0; t = time.perf_counter()
    if one_call:
        h.set_sample(sample, None, tree); T["set_sample (matrix + efflens + tree)"] = time.perf_counter() - t; t = time.perf_counter()
    else:
        h.set_sample(sample); T["set_matrix + set_efflens"] = time.perf_counter() - t; t = time.perf_counter()
        h.set_tree(*tree); T["set_tree"] = time.perf_counter() - t; t = time.perf_counter()
    h.init_params(); h.run_steps(500); T["enqueue 500 steps"] = time.perf_counter() - t; t = time.perf_counter()
    h.sync(); T["wait"] = time.perf_counter() - t; t = time.perf_counter()
    h.get_params(); T["get_params"] = time.perf_counter() - t; t = time.perf_counter()
    h.close(); T["destroy"] = time.perf_counter() - t
    print(rep, "one call   " if one_call else "three calls", "total %.3f" % (time.perf_counter() - t0), {k: round(v * 1e3, 1) for k, v in T.items()}, flush=True)
