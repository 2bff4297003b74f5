"""Run a few ADAM steps on C3 (or --config) without CUDA graphs so that ncu can list the per-kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
import polee_b200 as pb
from bench import generate

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
s, tree, K = generate(a.config, "cuda:0")
h = pb.Handle(num_mc_samples=K, num_steps=a.steps, use_cuda_graph=False)
cp, rv, nz = s["colptr"].to(torch.int32), s["rowval"].to(torch.int32), s["nzval"].contiguous()
h.set_matrix_device(s["m"], s["n"], cp.data_ptr(), rv.data_ptr(), nz.data_ptr())
h.set_efflens(s["efflens"].cpu().numpy())
h.set_tree(*tree)
h.init_params()
h.run_steps(a.steps)
h.sync()
print("done")
