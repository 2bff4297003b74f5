"""Stress the 'several handles on one GPU from several host threads' path (prep_many) and bisect with env switches."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import polee_b200 as pb
    from polee_b200 import synth
    samples, trees = [], []
    for i in range(5):
        s = synth.make_sample(40000 + 7000 * i, 2500, seed=300 + i)
        ns = synth.to_numpy_sample(s)
        samples.append(pb.RNASeqSample(ns["m"], ns["n"], ns["colptr"], ns["rowval"], ns["nzval"], ns["efflens"]))
        trees.append(synth.balanced_tree(2500, s["gene_sizes"].numpy()))
    seq = [pb.approximate_likelihood(pb.LogitSkewNormalPTTApprox(), s, tree_topology=t, num_steps=30, num_mc_samples=8)
           for s, t in zip(samples, trees)]
    bad = 0
    for rep in range(int(sys.argv[2])):
        try:
            par = pb.prep_many(samples, trees, devices=(0, 0, 0), num_steps=30, num_mc_samples=8)
            ok = all(np.array_equal(a[k], b[k]) for a, b in zip(seq, par) for k in ("mu", "omega", "alpha"))
            if not ok:
                bad += 1
                print("MISMATCH in rep", rep, flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("ERROR in rep", rep, repr(e)[:200], flush=True)
    print("bad", bad, "of", sys.argv[2])
    sys.exit(0)
for name, env in (("default", {}), ("no-cache", {"POLEE_NO_CACHE": "1"}), ("split", {"POLEE_LAYOUT": "split"}),
                  ("fused", {"POLEE_LAYOUT": "fused"})):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, __file__, "child", "12"], env=e, capture_output=True, text=True)
    print("==", name, "::", " | ".join(out.stdout.strip().splitlines()[-4:]), out.stderr.strip()[-300:])
