"""GPU parity of the three HSB ops against vectors produced by the reference's own hsb_ops.cpp."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["fixture_shared", "random97_shared", "per_row"])
def test_hsb_ops_match_reference_vectors(case):
    import polee_b200 as pb
    v = np.load(os.path.join(GOLDEN, "hsb_reference_vectors.npz"))
    g = lambda k: v["%s__%s" % (case, k)]  # noqa: E731
    L, R, F = g("left"), g("right"), g("leaf")
    for idx in ([L, R, F], [L[:1], R[:1], F[:1]] if case != "per_row" else None):
        if idx is None:
            continue
        x = pb.hsb(g("y_logit"), *idx)
        # y = 1/(1+exp(-y_logit)) uses the device's double exp(); everything else is the same IEEE sequence
        assert relerr(x, g("x")) <= 1e-6
        assert (x != g("x")).mean() < 0.02                        # in practice (almost) bit-identical
        y, ladj = pb.inv_hsb(g("x"), *idx)
        assert np.array_equal(y, g("y"))                           # add + divide only: bit-exact
        assert relerr(ladj, g("ladj")) <= 1e-6                     # Float32 accumulator replayed in reference order
        bp = pb.inv_hsb_grad(g("y_grad"), g("ladj_grad"), g("y"), g("ladj"), *idx)
        assert np.array_equal(bp, g("backprops"))                  # mul/add/div in the same association: bit-exact


def test_hsb_roundtrip_large_shared_tree(oracle):
    """HSB(logit(InvHSB(x))) = x on a 20 000-leaf tree, B = 16 (size-independent property)."""
    import polee_b200 as pb
    from polee_b200 import synth
    n, B = 20000, 16
    l, r, f = pb.make_inverse_ptt_params(*synth.balanced_tree(n))
    rng = np.random.default_rng(0)
    x = rng.dirichlet(np.ones(n) * 0.5, B).astype(np.float32).clip(1e-12)
    y, ladj = pb.inv_hsb(x, l, r, f)
    yo, ladj_o = oracle.inv_hsb(x[:2], l, r, f)
    assert np.array_equal(y[:2], yo) and relerr(ladj[:2], ladj_o) <= 1e-6
    yl = np.log(y / (1 - y)).astype(np.float32)
    x2 = pb.hsb(yl, l, r, f)
    s = x.astype(np.float64).sum(1, keepdims=True)
    np.testing.assert_allclose(x2, x / s, rtol=3e-5)


def test_hsb_bad_tree():
    import polee_b200 as pb
    with pytest.raises(pb.PoleeError):
        pb.hsb(np.zeros((1, 2), np.float32), [1, 3, -1, -1, -1], [2, 4, -1, -1, -1], [-1, -1, 0, 0, 1])


def test_tf_shim_through_stub_tensorflow():
    """The rebuilt TF plugin (polee_b200/tf/hsb_ops_b200.cpp: same REGISTER_OP signatures, Compute() forwards to the
    C ABI) driven through the stub TF API -- the same harness that runs the reference's own op file."""
    import ctypes as C
    from conftest import ROOT
    path = os.path.join(ROOT, "tests", "tf_shim", "libshim_hsb_ops_stubtf.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(path), "check"])
    import polee_b200  # noqa: F401  (loads libpolee_b200.so first)
    shim = C.CDLL(path)
    v = np.load(os.path.join(GOLDEN, "hsb_reference_vectors.npz"))
    P = C.c_void_p
    p = lambda a: a.ctypes.data_as(P)  # noqa: E731
    for case in ("fixture_shared", "per_row"):
        g = lambda k: np.ascontiguousarray(v["%s__%s" % (case, k)])  # noqa: E731
        L, R, F, yl = g("left"), g("right"), g("leaf"), g("y_logit")
        B, n = yl.shape[0], yl.shape[1] + 1
        x = np.zeros((B, n), np.float32)
        assert shim.shim_hsb(C.c_int64(B), C.c_int64(n), p(yl), p(L), p(R), p(F), p(x), C.c_int(1)) == 0
        assert relerr(x, g("x")) <= 1e-6
        y = np.zeros((B, n - 1), np.float64)
        ladj = np.zeros((B, 1), np.float32)
        xin = g("x")
        assert shim.shim_inv_hsb(C.c_int64(B), C.c_int64(n), p(xin), p(L), p(R), p(F), p(y), p(ladj), C.c_int(1)) == 0
        assert np.array_equal(y, g("y")) and relerr(ladj, g("ladj")) <= 1e-6
        bp = np.zeros((B, n), np.float32)
        yg, lg, yy, la = g("y_grad"), g("ladj_grad"), g("y"), g("ladj")
        assert shim.shim_inv_hsb_grad(C.c_int64(B), C.c_int64(n), p(yg), p(lg), p(yy), p(la), p(L), p(R), p(F), p(bp),
                                      C.c_int(1)) == 0
        assert np.array_equal(bp, g("backprops"))


def test_hsb_device_resident_entry_points():
    """polee_*_device: device tensors in and out on a caller's stream, no copies (what a DEVICE_GPU registration binds);
    identical to the host-tensor forms, shared and per-row trees."""
    import torch
    import polee_b200 as pb
    v = np.load(os.path.join(GOLDEN, "hsb_reference_vectors.npz"))
    st = torch.cuda.Stream()
    for case in ("fixture_shared", "per_row"):
        g = lambda k: np.ascontiguousarray(v["%s__%s" % (case, k)])  # noqa: E731
        yl = g("y_logit")
        B, n = yl.shape[0], yl.shape[1] + 1
        idx = [g("left"), g("right"), g("leaf")]
        if case == "fixture_shared":
            idx = [a[:1] for a in idx]
        plan = pb.HsbPlan(n, *idx)
        dev = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
        with torch.cuda.stream(st):
            d_yl, d_x = dev(yl), torch.empty((B, n), dtype=torch.float32, device="cuda")
            d_xin = dev(g("x"))
            d_y = torch.empty((B, n - 1), dtype=torch.float64, device="cuda")
            d_ladj = torch.empty((B, 1), dtype=torch.float32, device="cuda")
            d_yg, d_lg, d_yy = dev(g("y_grad")), dev(g("ladj_grad")), dev(g("y"))
            d_bp = torch.empty((B, n), dtype=torch.float32, device="cuda")
            for rep in range(2):                                     # the second call reuses the plan's scratch
                plan.hsb_device(B, d_yl.data_ptr(), d_x.data_ptr(), st.cuda_stream)
                plan.inv_hsb_device(B, d_xin.data_ptr(), d_y.data_ptr(), d_ladj.data_ptr(), st.cuda_stream)
                plan.inv_hsb_grad_device(B, d_yg.data_ptr(), d_lg.data_ptr(), d_yy.data_ptr(), d_bp.data_ptr(), st.cuda_stream)
        st.synchronize()
        assert np.array_equal(d_x.cpu().numpy(), pb.hsb(yl, *idx))
        y_h, ladj_h = pb.inv_hsb(g("x"), *idx)
        assert np.array_equal(d_y.cpu().numpy(), y_h) and np.array_equal(d_ladj.cpu().numpy(), ladj_h)
        assert np.array_equal(d_y.cpu().numpy(), g("y")) and relerr(d_ladj.cpu().numpy(), g("ladj")) <= 1e-6
        assert np.array_equal(d_bp.cpu().numpy(), g("backprops"))
        plan.close()


def test_tf_shim_gpu_registration_through_stub_tensorflow():
    """The shim's DEVICE_GPU kernels (index tensors in HostMemory, data tensors consumed and produced in device memory on
    the op's stream), driven through the stub TF API with torch-owned device buffers."""
    import ctypes as C
    import torch
    from conftest import ROOT
    path = os.path.join(ROOT, "tests", "tf_shim", "libshim_hsb_ops_stubtf.so")
    import polee_b200  # noqa: F401
    shim = C.CDLL(path)
    v = np.load(os.path.join(GOLDEN, "hsb_reference_vectors.npz"))
    P = C.c_void_p
    p = lambda a: a.ctypes.data_as(P)  # noqa: E731
    d = lambda t: P(t.data_ptr())  # noqa: E731
    st = torch.cuda.Stream()
    for case in ("fixture_shared", "per_row"):
        g = lambda k: np.ascontiguousarray(v["%s__%s" % (case, k)])  # noqa: E731
        L, R, F, yl = g("left"), g("right"), g("leaf"), g("y_logit")
        B, n = yl.shape[0], yl.shape[1] + 1
        with torch.cuda.stream(st):
            t_yl, t_x = torch.from_numpy(yl).cuda(), torch.zeros((B, n), dtype=torch.float32, device="cuda")
            assert shim.shim_hsb_gpu(C.c_int64(B), C.c_int64(n), d(t_yl), p(L), p(R), p(F), d(t_x), P(st.cuda_stream)) == 0
            t_xin = torch.from_numpy(g("x")).cuda()
            t_y = torch.zeros((B, n - 1), dtype=torch.float64, device="cuda")
            t_ladj = torch.zeros((B, 1), dtype=torch.float32, device="cuda")
            assert shim.shim_inv_hsb_gpu(C.c_int64(B), C.c_int64(n), d(t_xin), p(L), p(R), p(F), d(t_y), d(t_ladj),
                                         P(st.cuda_stream)) == 0
            t_yg, t_lg = torch.from_numpy(g("y_grad")).cuda(), torch.from_numpy(g("ladj_grad")).cuda()
            t_yy, t_la = torch.from_numpy(g("y")).cuda(), torch.from_numpy(g("ladj")).cuda()
            t_bp = torch.zeros((B, n), dtype=torch.float32, device="cuda")
            assert shim.shim_inv_hsb_grad_gpu(C.c_int64(B), C.c_int64(n), d(t_yg), d(t_lg), d(t_yy), d(t_la), p(L), p(R), p(F),
                                              d(t_bp), P(st.cuda_stream)) == 0
        st.synchronize()
        assert relerr(t_x.cpu().numpy(), g("x")) <= 1e-6
        assert np.array_equal(t_y.cpu().numpy(), g("y")) and relerr(t_ladj.cpu().numpy(), g("ladj")) <= 1e-6
        assert np.array_equal(t_bp.cpu().numpy(), g("backprops"))
