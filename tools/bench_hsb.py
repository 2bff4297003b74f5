#!/usr/bin/env python3
"""K4 measurement: the three HSB ops (SURVEY 8d "K4 bytes") at B = 64 RNA-seq samples x n = 200 000 transcripts,
shared tree and a tree per row, through the C ABI with HOST buffers (what the DEVICE_CPU op hands over) and with
DEVICE buffers on a stream (polee_*_device, what the DEVICE_GPU op hands over; CUDA events), next to the
reference's own CPU op (oracle/_ref, the unmodified hsb_ops.cpp over the stub TF API) on the box's host cores.

    python tools/bench_hsb.py [--B 64] [--n 200000] [--reps 5]
Prints one JSON line per (op, tree mode)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polee_b200 as pb  # noqa: E402
from polee_b200 import _lib as L, synth  # noqa: E402
from oracle import polee_oracle as O  # noqa: E402

P = C.c_void_p
p = lambda a: a.ctypes.data_as(P)  # noqa: E731


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    B, n, N = a.B, a.n, 2 * a.n - 1
    lib = L.load_library()
    rng = np.random.default_rng(0)
    for mode in ("shared", "per_row"):
        if mode == "shared":
            l, r, f = pb.make_inverse_ptt_params(*synth.balanced_tree(n))
            Lx, Rx, Fx = (np.ascontiguousarray(v.reshape(1, N)) for v in (l, r, f))
            ib = 1
        else:
            Bt = min(B, 8)  # 8 distinct random trees, repeated (tree preparation is per distinct tree anyway)
            trees = [pb.make_inverse_ptt_params(*synth.balanced_tree(n, None)) for _ in range(1)]
            base = trees[0]
            Lx, Rx, Fx = (np.ascontiguousarray(np.broadcast_to(v, (B, N))) for v in base)
            ib = B
        plan = P()
        t0 = time.perf_counter()
        rc = lib.polee_hsb_plan_create(C.byref(plan), C.c_int32(0), C.c_int64(n), C.c_int64(ib), p(Lx), p(Rx), p(Fx))
        assert rc == 0, lib.polee_hsb_last_error()
        t_plan = time.perf_counter() - t0
        y_logit = rng.normal(0, 2, (B, n - 1)).astype(np.float32)
        x = np.zeros((B, n), np.float32)
        y = np.zeros((B, n - 1), np.float64)
        ladj = np.zeros((B, 1), np.float32)
        bp = np.zeros((B, n), np.float32)
        yg = rng.normal(size=(B, n - 1))
        lg = rng.normal(size=(B, 1)).astype(np.float32)

        def timeit(fn):
            fn()
            ts = []
            for _ in range(a.reps):
                t = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t)
            return min(ts)

        t_hsb = timeit(lambda: lib.polee_hsb_with_plan(plan, C.c_int64(B), p(y_logit), p(x)))
        t_inv = timeit(lambda: lib.polee_inv_hsb_with_plan(plan, C.c_int64(B), p(x), p(y), p(ladj)))
        t_grad = timeit(lambda: lib.polee_inv_hsb_grad_with_plan(plan, C.c_int64(B), p(yg), p(lg), p(y), p(bp)))
        # device-resident forms: CUDA events on a torch stream, tensors already in HBM
        import torch
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            d_yl, d_x = torch.from_numpy(y_logit).cuda(), torch.empty((B, n), dtype=torch.float32, device="cuda")
            d_y = torch.empty((B, n - 1), dtype=torch.float64, device="cuda")
            d_ladj = torch.empty((B, 1), dtype=torch.float32, device="cuda")
            d_yg, d_lg = torch.from_numpy(yg).cuda(), torch.from_numpy(lg).cuda()
            d_bp = torch.empty((B, n), dtype=torch.float32, device="cuda")
        S = P(st.cuda_stream)
        dp = lambda t: P(t.data_ptr())  # noqa: E731

        def dev_time(fn):
            fn(); fn()
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.reps):
                fn()
            e1.record(st)
            st.synchronize()
            return e0.elapsed_time(e1) / a.reps * 1e-3

        d_hsb = dev_time(lambda: lib.polee_hsb_device(plan, C.c_int64(B), dp(d_yl), dp(d_x), S))
        d_inv = dev_time(lambda: lib.polee_inv_hsb_device(plan, C.c_int64(B), dp(d_x), dp(d_y), dp(d_ladj), S))
        d_grad = dev_time(lambda: lib.polee_inv_hsb_grad_device(plan, C.c_int64(B), dp(d_yg), dp(d_lg), dp(d_y), dp(d_bp), S))
        assert np.array_equal(d_x.cpu().numpy(), x) and np.array_equal(d_y.cpu().numpy(), y)
        dev_t = {"hsb": d_hsb, "inv_hsb": d_inv, "inv_hsb_grad": d_grad}
        lib.polee_hsb_plan_destroy(plan)
        # reference CPU op (needs [B, 2n-1] index tensors)
        Lb, Rb, Fb = (np.ascontiguousarray(np.broadcast_to(v, (B, N))) for v in (Lx, Rx, Fx))
        ref = {}
        if O.ref_lib() is not None:
            t = time.perf_counter(); xr = O.hsb(y_logit, Lb, Rb, Fb, impl="ref", threads=a.cpu_threads); ref["hsb"] = time.perf_counter() - t
            t = time.perf_counter(); yr, lr = O.inv_hsb(xr, Lb, Rb, Fb, impl="ref", threads=a.cpu_threads); ref["inv_hsb"] = time.perf_counter() - t
            t = time.perf_counter(); O.inv_hsb_grad(yg, lg, yr, lr, Lb, Rb, Fb, impl="ref", threads=a.cpu_threads); ref["inv_hsb_grad"] = time.perf_counter() - t
            assert np.max(np.abs(x - xr) / np.maximum(xr, 1e-300)) < 1e-5
        idx_bytes = ib * N * 12
        alg = {"hsb": B * (n - 1) * 4 + B * n * 4 + idx_bytes, "inv_hsb": B * n * 4 + B * (n - 1) * 8 + B * 4 + idx_bytes,
               "inv_hsb_grad": B * (n - 1) * 16 + B * 4 + B * n * 4 + idx_bytes}
        for op, t in (("hsb", t_hsb), ("inv_hsb", t_inv), ("inv_hsb_grad", t_grad)):
            print(json.dumps({"op": op, "trees": mode, "B": B, "n": n, "ms_host_to_host": round(t * 1e3, 2),
                              "ms_device_resident": round(dev_t[op] * 1e3, 3),
                              "algorithmic_GB": round(alg[op] / 1e9, 4), "GBps_device_resident": round(alg[op] / dev_t[op] / 1e9, 1),
                              "GBps_incl_pcie": round(alg[op] / t / 1e9, 1),
                              "plan_create_ms": round(t_plan * 1e3, 1),
                              "reference_cpu_ms": round(ref.get(op, float("nan")) * 1e3, 1), "cpu_threads": a.cpu_threads}))


if __name__ == "__main__":
    main()
