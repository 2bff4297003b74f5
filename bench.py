#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through libpolee_b200.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Metric: ELBO-gradient evaluations per second (one eval = one Monte-Carlo draw's full forward+backward:
K1 SpMM + K2 transposed gradient + K3 reparameterisation/tree fwd+bwd/ADAM) on the synthetic GENCODE-scale
config C3: 30 M fragments x 200 k transcripts, nnz ~ 120 M, K = 8 draws per ADAM step.  A "step" is one ADAM
step = K evals.  At N > 1 the SAME matrix is row-partitioned into equal-nnz blocks (strong scaling) and the
transcript-length gradient is all-reduced once per step over NCCL.

value   : steady-state evals/s with the matrix resident in HBM (CUDA events on the handle's stream, max over ranks)
e2e     : the same metric through the public API with HOST buffers: one whole approximate_likelihood call
          (upload of the CSC arrays from pinned host memory, device-side layout conversion, 500 ADAM steps,
          download of mu/omega/alpha), evals / wall seconds
roofline: dominant sparse kernel, algorithmic bytes (SURVEY 8d formula) / CUDA-event time vs the measured HBM peak
cpu_baseline: the oracle (C restatement of the reference's multithreaded Julia loops) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (m, n, K, long_rows, seed)
    "c2": (1 << 20, 20_000, 1, False, 20260002),
    "c3": (30_000_000, 200_000, 8, False, 20260003),
    "c3-small": (3_000_000, 200_000, 8, False, 20260003),
}
FIT_STEPS = 500  # LIKAP_NUM_STEPS (src/constants.jl:64): what one approximate_likelihood call runs


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region, sampled through NVML every few ms (the nvidia-smi
    query of B200_PROFILING.md takes longer than a whole timed region here); falls back to nvidia-smi."""

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu = gpu_index
        self.period = period_s
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self.stop = threading.Event()
        self.t = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run(self):
        n = self.nvml
        bits = {}
        if n is not None:
            for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
                v = getattr(n, attr, None) or getattr(n, attr.replace("ClocksEventReason", "ClocksThrottleReason"), None)
                if v is not None:
                    bits[name] = v
        while not self.stop.is_set():
            try:
                if n is not None:
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                    get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                    r = get(self.h)
                    for name, b in bits.items():
                        if r & b:
                            self.reasons.add(name)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = out.strip().split(",")
                    self.sm.append(float(a)); self.sm_max = float(b)
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def generate(cfg, device):
    """Synthetic sample (polee-synth-v1) as torch tensors on `device` + host tree arrays."""
    import torch
    from polee_b200 import synth
    m, n, K, long_rows, seed = CONFIGS[cfg]
    s = synth.make_sample(m, n, seed=seed, device=device, long_rows=long_rows)
    tree = synth.balanced_tree(n, s["gene_sizes"].cpu().numpy())
    return s, tree, K


def row_block_device(s, lo, hi):
    """CSC of rows [lo, hi) (0-based) of the device-resident sample, all n columns; int32 index tensors."""
    import torch
    n = s["n"]
    rowval, nzval = s["rowval"], s["nzval"]
    if lo == 0 and hi == s["m"]:
        return hi - lo, s["colptr"].to(torch.int32), rowval.to(torch.int32), nzval
    keep = (rowval > lo) & (rowval <= hi)
    counts = torch.diff(s["colptr"])
    col_of = torch.repeat_interleave(torch.arange(n, device=rowval.device), counts)
    cnt = torch.bincount(col_of[keep], minlength=n)
    colptr = torch.cat([torch.ones(1, dtype=torch.int64, device=rowval.device), 1 + torch.cumsum(cnt, 0)])
    return hi - lo, colptr.to(torch.int32), (rowval[keep] - lo).to(torch.int32), nzval[keep].contiguous()


def equal_nnz_bounds(s, parts):
    import torch
    rows = torch.bincount(s["rowval"] - 1, minlength=s["m"])
    cum = torch.cumsum(rows, 0)
    nnz = int(cum[-1].item())
    b = [0]
    for p in range(1, parts):
        b.append(int(torch.searchsorted(cum, torch.tensor([nnz * p // parts], device=cum.device)).item()))
    b.append(s["m"])
    return b


def run_ours(args):
    import torch
    import torch.distributed as dist
    import polee_b200 as pb
    from polee_b200 import api as pbapi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    s, tree, K = generate(args.config, dev)
    m, n = s["m"], s["n"]
    efflens = s["efflens"].cpu().numpy()
    nnz_total = s["nnz"]
    bounds = equal_nnz_bounds(s, world)
    m_loc, colptr_d, rowval_d, nzval_d = row_block_device(s, bounds[rank], bounds[rank + 1])
    nnz_loc = int(rowval_d.numel())

    # pinned host copies for the e2e arm: the whole matrix at N = 1 (also feeds the CPU baseline), each rank's own
    # row block at N > 1
    host = None
    pin = lambda t, dt: torch.empty(t.shape, dtype=dt, pin_memory=True).copy_(t.to(dt)).numpy()  # noqa: E731
    if not args.no_e2e:
        if world == 1:
            host = {"colptr": pin(s["colptr"], torch.int32).view(np.uint32), "rowval": pin(s["rowval"], torch.int32).view(np.uint32),
                    "nzval": pin(s["nzval"], torch.float32)}
        else:
            host = {"colptr": pin(colptr_d, torch.int32).view(np.uint32), "rowval": pin(rowval_d, torch.int32).view(np.uint32),
                    "nzval": pin(nzval_d, torch.float32)}
    del s
    torch.cuda.empty_cache()

    h = pb.Handle(device=local, num_mc_samples=K, num_steps=max(args.steps + args.warmup, 1), seed=args.seed)
    h.set_matrix_device(m_loc, n, colptr_d.data_ptr(), rowval_d.data_ptr(), nzval_d.data_ptr())
    del colptr_d, rowval_d, nzval_d
    torch.cuda.empty_cache()
    h.set_efflens(efflens)
    h.set_tree(*tree)
    if world > 1:
        uid = [pbapi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(world, rank, uid[0])
    stats = h.step_stats()

    stream = torch.cuda.ExternalStream(h.stream(), device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: K steps, device time on the launching stream, max over ranks
    h.init_params()
    h.run_steps(args.warmup)
    h.sync()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        with torch.cuda.stream(stream):
            e0.record(stream)
        h.run_steps(args.steps)
        with torch.cuda.stream(stream):
            e1.record(stream)
        h.sync()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = clk.summary()
    evals_per_s = K * args.steps / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events on the handle's stream, back-to-back launches; inputs >> L2)
    reps = max(5, min(50, args.steps))
    t_k1, t_k2, t_k3 = (h.time_kernel(w, reps) for w in (1, 2, 3))
    peak, peak_src = measured_peak()
    # algorithmic bytes of THIS rank's launches (SURVEY 8d formula, local nnz / rows, padded draw count KP)
    KP = 1
    while KP < K:
        KP *= 2
    b_k1 = nnz_loc * 8 + (m_loc + 1) * 4 + KP * n * 4 + KP * m_loc * 4
    b_k2 = nnz_loc * 8 + (n + 1) * 4 + KP * m_loc * 4 + KP * n * 4
    fused = stats["bytes_k2"] == 0      # this rank's block uses the fused row-tile layout (one sparse pass per step)
    if fused:
        # ONE launch does the work of K1 and K2, so its algorithmic bytes are SURVEY 8(d)'s B_K1 + B_K2 (matrix twice +
        # the w round trip); what it really moves is far less (`traffic`, `moved_gbs`: w never leaves the SM) -- it is
        # bound by the shared-memory pipe and the CTA barriers, not by HBM (profiles/README.md)
        b_f = stats["bytes_k1"]
        achieved = (b_k1 + b_k2) / (t_k1 * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k12_fused_rowtiles", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                    "note": "achieved = algorithmic bytes of K1 + K2 (SURVEY 8d) / time of the one fused pass incl. its "
                            "second stage; moved_gbs = bytes the pass actually streams / same time",
                    "kernels_ms": {"k12_fused_rowtiles": round(t_k1, 4), "k3_tree_reparam_adam": round(t_k3, 4)},
                    "kernels_gbs": {"k12_fused_rowtiles": round(achieved, 1)},
                    "moved_gbs": round(b_f / t_k1 / 1e6, 1),
                    "step_gbs": round((b_k1 + b_k2 + stats["bytes_k3"]) / (ms / args.steps) / 1e6, 1)}
        dom = ("k12_fused_rowtiles", b_f, t_k1)
    else:
        dom = ("k1_sell_fwd", b_k1, t_k1) if t_k1 >= t_k2 else ("k2_csc_grad", b_k2, t_k2)
        achieved = dom[1] / (dom[2] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                    "kernels_ms": {"k1_sell_fwd": round(t_k1, 4), "k2_csc_grad": round(t_k2, 4), "k3_tree_reparam_adam": round(t_k3, 4)},
                    "kernels_gbs": {"k1_sell_fwd": round(b_k1 / t_k1 / 1e6, 1), "k2_csc_grad": round(b_k2 / t_k2 / 1e6, 1)},
                    "step_gbs": round((b_k1 + b_k2 + stats["bytes_k3"]) / (ms / args.steps) / 1e6, 1)}
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.config)
    if os.path.exists(tr) and world == 1:  # the capture is of the whole-sample launch
        try:
            roofline["traffic"] = json.load(open(tr)).get(dom[0])
        except Exception:
            pass

    # ---- e2e: one whole fit through the public API from HOST buffers
    e2e = None
    cpu = None
    if host is not None and world == 1:
        h.close()
        torch.cuda.empty_cache()
        sample = pb.RNASeqSample(m, n, host["colptr"], host["rowval"], host["nzval"], efflens)
        for rep in range(2):   # first call warms the CUDA context / allocator; the second is reported
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = pb.approximate_likelihood(pb.LogitSkewNormalPTTApprox(), sample, tree_topology=tree, num_steps=FIT_STEPS,
                                            num_mc_samples=K, seed=args.seed, device=local)
            t_fit = time.perf_counter() - t0
        assert np.all(np.isfinite(out["mu"]))
        h2d = host["colptr"].nbytes + host["rowval"].nbytes + host["nzval"].nbytes + efflens.nbytes + 2 * 4 * (2 * n - 1)
        e2e = {"value": round(K * FIT_STEPS / t_fit, 1), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(3 * 4 * (n - 1)), "fit_time_s": round(t_fit, 4), "adam_steps_per_fit": FIT_STEPS,
               "step": "one approximate_likelihood call: CSC upload from pinned host memory + device layout build + "
                       "%d ADAM steps x %d draws + parameter download" % (FIT_STEPS, K)}
        if not args.no_cpu:
            cpu = cpu_baseline(m, n, K, host, efflens, tree, budget_s=args.cpu_budget)
    elif host is not None:
        # N > 1: every rank re-feeds its own row block from host memory into its (comm-initialised) handle and runs
        # the whole 500-step fit; wall time, max over ranks
        block = pb.RNASeqSample(m_loc, n, host["colptr"], host["rowval"], host["nzval"], efflens)
        h.opts.num_steps = FIT_STEPS
        t_fit = None
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            h.set_sample(block)
            h.set_tree(*tree)
            h.init_params()
            h.run_steps(FIT_STEPS)
            h.sync()
            out = h.get_params()
            tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_fit = float(tt.item())
        h2d = host["colptr"].nbytes + host["rowval"].nbytes + host["nzval"].nbytes + efflens.nbytes + 2 * 4 * (2 * n - 1)
        hb = torch.tensor([h2d], device=dev, dtype=torch.float64)
        dist.all_reduce(hb)
        e2e = {"value": round(K * FIT_STEPS / t_fit, 1), "unit": "evals/s", "h2d_bytes_per_step": int(hb.item()),
               "d2h_bytes_per_step": int(3 * 4 * (n - 1)), "fit_time_s": round(t_fit, 4), "adam_steps_per_fit": FIT_STEPS,
               "step": "every rank: its row block's CSC upload from pinned host memory + layout build + %d ADAM steps x %d "
                       "draws (one all-reduce each) + parameter download; wall time, max over ranks" % (FIT_STEPS, K)}
        h.close()
    else:
        h.close()
    torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": "elbo_grad_evals_per_sec", "value": round(evals_per_s, 1), "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 storage, f64 accumulate",
                "data": "synthetic (polee-synth-v1, seed %d)" % CONFIGS[args.config][4],
                "config": {"workload": "%s: %d fragments x %d transcripts, nnz %d, K=%d draws/step, balanced-by-gene tree; "
                                       "row-partitioned into %d equal-nnz block(s)" % (args.config, m, n, nnz_total, K, world),
                           "l2": "inputs (%.1f GB/step streamed) are far larger than the 126 MB L2; no flush needed"
                                 % ((stats["bytes_k1"] if fused else b_k1 + b_k2) / 1e9),
                           "layout": "fused row tiles (one sparse pass per step)" if fused else "split (SELL slabs for K1 + re-sorted CSC for K2)",
                           "noise": "device Philox"},
                "clocks": clocks, "gpu_launches": int(stats["launches"] * args.steps), "roofline": roofline}
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(m, n, K, host, efflens, tree, budget_s=25.0, warm=0):
    """The oracle (C/OpenMP restatement of the reference's Julia loops) on the host cores: whole ADAM steps of
    K draws on the SAME matrix, as many as fit in the budget (at least one)."""
    from oracle import polee_oracle as O
    st = O.FitStepper(m, n, host["colptr"], host["rowval"], host["nzval"], efflens, tree[0], tree[1], num_mc_samples=K)
    for _ in range(warm):
        st.step()
    t0 = time.perf_counter()
    steps = 0
    while True:
        st.step()
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or el + el / steps > budget_s * 1.3:
            break
    st.close()
    return {"value": round(K * steps / el, 3), "unit": "evals/s", "cores": O.num_threads(), "kind": "port",
            "sample": "%d full ADAM step(s) of %d draws on the same %d x %d matrix (setup/transposition excluded), "
                      "%.1f s" % (steps, K, m, n, el)}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port -- Julia is not installable here) with all host
    threads, same config/metric.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from polee_b200 import synth
    m, n, K, long_rows, seed = CONFIGS[args.config]
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    s = synth.make_sample(m, n, seed=seed, device=dev, long_rows=long_rows)
    tree = synth.balanced_tree(n, s["gene_sizes"].cpu().numpy())
    ns = synth.to_numpy_sample(s)
    nnz = s["nnz"]
    del s
    from oracle import polee_oracle as O
    st = O.FitStepper(m, n, ns["colptr"], ns["rowval"], ns["nzval"], ns["efflens"], tree[0], tree[1], num_mc_samples=K)
    budget = args.cpu_budget * 4
    t_w = time.perf_counter()
    done_w = 0
    for _ in range(args.warmup):
        st.step()
        done_w += 1
        if time.perf_counter() - t_w > budget / 4:
            break
    t0 = time.perf_counter()
    steps = 0
    for _ in range(args.steps):
        st.step()
        steps += 1
        if time.perf_counter() - t0 > budget:
            break
    el = time.perf_counter() - t0
    st.close()
    v = round(K * steps / el, 3)
    cores = O.num_threads()
    sample = "%d of the %d requested ADAM steps (x %d draws) on the full matrix within a %.0f s budget; %d warm-up" % (
        steps, args.steps, K, budget, done_w)
    print(json.dumps({
        "impl": "reference", "metric": "elbo_grad_evals_per_sec", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": done_w, "ms_per_step": round(el / steps * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 storage, f64 accumulate", "data": "synthetic (polee-synth-v1, seed %d)" % seed,
        "config": {"workload": "%s: %d fragments x %d transcripts, nnz %d, K=%d draws/step, balanced-by-gene tree" % (
            args.config, m, n, nnz, K), "note": "CPU restatement of the reference's multithreaded Julia path "
            "(oracle/polee_oracle.c, OpenMP static chunks = Threads.@threads); Julia itself is not available"},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--seed", type=int, default=123456789)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
