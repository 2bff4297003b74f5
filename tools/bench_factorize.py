"""Exact row de-duplication (polee_exact_factorization): host arrays in, host arrays out.
Times the device implementation on C2 / C3-shaped samples and the oracle (Python dict over rows, like the reference's
Julia Dict) on C2.  Synthetic rows carry i.i.d. values, so few rows collapse; every 4th row is therefore repeated."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import polee_b200 as pb
from polee_b200 import synth
from bench import generate

for cfg in sys.argv[1:] or ["c2", "c3"]:
    s, tree, K = generate(cfg, "cuda:0")
    ns = synth.to_numpy_sample(s)
    del s
    torch.cuda.empty_cache()
    sample = pb.RNASeqSample(ns["m"], ns["n"], ns["colptr"], ns["rowval"], ns["nzval"], ns["efflens"])
    for rep in range(2):
        t0 = time.perf_counter()
        comp, counts = pb.exact_factorization(sample)
        dt = time.perf_counter() - t0
    print("%s: m = %d, nnz = %d -> %d unique rows in %.3f s (%.1f M rows/s, host to host)"
          % (cfg, sample.m, len(sample.rowval), comp.m, dt, sample.m / dt / 1e6), flush=True)
    if sample.m <= 2_000_000:
        from oracle import polee_oracle as O
        t0 = time.perf_counter()
        mu, *_ = O.exact_factorization(sample.m, sample.n, sample.colptr, sample.rowval, sample.nzval)
        dt = time.perf_counter() - t0
        print("%s: oracle (1 host thread, dict of rows) %d unique rows in %.2f s (%.2f M rows/s)" % (cfg, mu, dt, sample.m / dt / 1e6))
        assert mu == comp.m
