"""Time the fused likelihood pass (kernel + second stage) on one generated sample for several tile geometries.
usage: python tools/fused_sweep.py [--config c3] rows:window[:ctas] ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
import polee_b200 as pb
from bench import generate

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--ranks", type=int, default=1, help="time rank 0's row block of an R-way equal-nnz partition")
ap.add_argument("geoms", nargs="*", default=["512:4096"])
a = ap.parse_args()
s, tree, K = generate(a.config, "cuda:0")
from bench import row_block_device, equal_nnz_bounds
m_loc = s["m"]
if a.ranks > 1:
    bounds = equal_nnz_bounds(s, a.ranks)
    m_loc, cp, rv, nz = row_block_device(s, bounds[0], bounds[1])
    nz = nz.contiguous()
else:
    cp, rv, nz = s["colptr"].to(torch.int32), s["rowval"].to(torch.int32), s["nzval"].contiguous()
eff = s["efflens"].cpu().numpy()
for geom in a.geoms:
    parts = geom.split(":")
    if parts[0] == "split":
        os.environ["POLEE_LAYOUT"] = "split"
    else:
        os.environ["POLEE_LAYOUT"] = "fused"
        os.environ["POLEE_FT_ROWS"], os.environ["POLEE_FT_WINDOW"] = parts[0], parts[1]
        if len(parts) > 2:
            os.environ["POLEE_FUSED_CTAS"] = parts[2]
        else:
            os.environ.pop("POLEE_FUSED_CTAS", None)
    h = pb.Handle(num_mc_samples=K, num_steps=10)
    h.set_matrix_device(m_loc, s["n"], cp.data_ptr(), rv.data_ptr(), nz.data_ptr())
    h.set_efflens(eff)
    h.set_tree(*tree)
    h.init_params()
    h.run_steps(3)
    h.sync()
    t1 = h.time_kernel(1, 20)
    t2 = h.time_kernel(2, 20)
    st = h.step_stats()
    print("%-16s likelihood pass %.4f ms (+ %.4f)  bytes %.0f MB" % (geom, t1, t2, st["bytes_k1"] / 1e6), flush=True)
    h.close()
