"""BASELINE config 5 in miniature: whole samples sharded over GPUs ("replicas only": one handle per device, a host
work queue, no collective -- the sample loop of `polee prep`, src/main.jl:590-631).
usage: python tools/bench_prep_many.py [--samples 16] [--config c3-small] [--devices 0,1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import polee_b200 as pb
from polee_b200 import synth
from bench import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=16)
ap.add_argument("--config", default="c3-small")
ap.add_argument("--devices", default="0")
a = ap.parse_args()
devices = tuple(int(d) for d in a.devices.split(","))
m, n = CONFIGS[a.config][0], CONFIGS[a.config][1]
rng = np.random.default_rng(5)
samples, trees = [], []
for i in range(a.samples):
    mi = int(m * np.exp(rng.normal(0, 0.35)))                       # depth spread around the config's m
    s = synth.make_sample(mi, n, seed=500 + i, device="cuda:0")
    ns = synth.to_numpy_sample(s)
    samples.append(pb.RNASeqSample(ns["m"], ns["n"], ns["colptr"], ns["rowval"], ns["nzval"], ns["efflens"]))
    trees.append(synth.balanced_tree(n, s["gene_sizes"].cpu().numpy()))
    del s
torch.cuda.empty_cache()
for rep in range(2):
    t0 = time.perf_counter()
    out = pb.prep_many(samples, trees, devices=devices, num_steps=500, num_mc_samples=8)
    dt = time.perf_counter() - t0
rows = sum(s.m for s in samples)
print("%d samples (%s shape, %.1f M fragments in total) on %d GPU(s): %.2f s = %.1f samples/s, %.0f ELBO-grad evals/s aggregate"
      % (len(samples), a.config, rows / 1e6, len(devices), dt, len(samples) / dt, len(samples) * 500 * 8 / dt))
assert all(np.isfinite(o["mu"]).all() for o in out)
