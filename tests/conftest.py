import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


def _have_gpu():
    """True when libpolee_b200 sees a CUDA device (gpu-marked tests are skipped only when there is none at all, e.g. a
    plain `pytest tests` on the CPU-only build box; on any GPU box they run and fail loudly if something is wrong)."""
    try:
        import ctypes as C
        from polee_b200 import _lib as L
        lib = L.load_library()
        cc = (C.c_int32 * 2)()
        sms = C.c_int32()
        mem = C.c_int64()
        rc = lib.polee_device_info(C.c_int32(0), cc, C.byref(sms), C.byref(mem))
        return rc == 0 and cc[0] > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if any("gpu" in it.keywords for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device visible to libpolee_b200")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


class Fixture:
    """The reference's one golden pair: likelihood-matrix.h5 (input) and prep.h5 (tree + fit)."""

    def __init__(self):
        f = lambda name, dt: np.fromfile(os.path.join(GOLDEN, name), dt)  # noqa: E731
        self.m, self.n = 19743, 313
        self.colptr = f("fixture_colptr.u32", np.uint32)
        self.rowval = f("fixture_rowval.u32", np.uint32)
        self.nzval = f("fixture_nzval.f32", np.float32)
        self.efflens = f("fixture_efflens.f32", np.float32)
        self.parent_idxs = f("fixture_prep_node_parent_idxs.i32", np.int32)
        self.js = f("fixture_prep_node_js.i32", np.int32)
        self.mu = f("fixture_prep_mu.f32", np.float32)
        self.omega = f("fixture_prep_omega.f32", np.float32)
        self.alpha = f("fixture_prep_alpha.f32", np.float32)


@pytest.fixture(scope="session")
def fx():
    return Fixture()


@pytest.fixture(scope="session")
def oracle():
    from oracle import polee_oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def small_synth():
    """m = 20 000, n = 1 500 synthetic sample + a balanced-by-gene tree (CPU generated, seeded)."""
    from polee_b200 import synth
    s = synth.make_sample(20000, 1500, seed=11)
    ns = synth.to_numpy_sample(s)
    ns["tree"] = synth.balanced_tree(1500, s["gene_sizes"].numpy())
    return ns


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def sample_loglik(O, fx, mu, omega, alpha, ndraws=400, seed=0):
    """ApproxLikelihoodSampler.rand! (src/approx-sampler.jl:37-44) restated with the oracle's pieces: mean log
    likelihood of draws from the fitted approximation, and the posterior mean of x."""
    rng = np.random.default_rng(seed)
    t = O.PTT(fx.parent_idxs, fx.js)
    M = O.Model(fx.m, fx.n, fx.colptr, fx.rowval, fx.nzval)
    sigma = np.exp(omega.astype(np.float64))
    lps, xmean = [], np.zeros(fx.n)
    for _ in range(ndraws):
        z0 = rng.normal(size=fx.n - 1)
        z = np.sinh(np.arcsinh(z0) + alpha)
        y = 1.0 / (1.0 + np.exp(-(mu + sigma * z)))
        y = np.clip(y, 1e-10, 1 - 1e-10)
        x, _ = t.transform(y)
        x = np.maximum(x, np.float32(1e-10))
        lp, _ = M.log_likelihood(x, gradonly=False)
        lps.append(lp)
        xmean += x
    return float(np.mean(lps)), xmean / ndraws
