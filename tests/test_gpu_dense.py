"""Dense heads on the device: the fixture-geomodel KAT and bat-head shapes."""
import numpy as np
import pytest

import birda_b200 as b

pytestmark = pytest.mark.gpu

# tests/fixtures/make_fixture_geomodel.py:20-28 (weights and bias are in the reference's source)
W_FIX = np.array([[0.010, -0.020, 0.030, 0.001, 0.050], [0.005, 0.010, -0.015, 0.002, 0.020],
                  [0.100, 0.050, -0.200, 0.010, 0.150]], np.float32)
B_FIX = np.array([0.5, -3.0, 0.2, -9.0, 1.0], np.float32)


def run(ctx, x, W, bias, act):
    import torch
    dx, dW = torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda()
    db = torch.from_numpy(bias).cuda() if bias is not None else None
    out = torch.empty((x.shape[0], W.shape[1]), dtype=torch.float32, device="cuda")
    ctx.dense_run(dx.data_ptr(), x.shape[0], x.shape[1], dW.data_ptr(), db.data_ptr() if db is not None else None,
                  W.shape[1], act, out.data_ptr())
    ctx.sync()
    return out.cpu().numpy()


def test_fixture_geomodel_kat():
    """sigmoid([lat, lon, week] W + B) at Helsinki (tests/geomodel_range_filter.rs:33-37): species 3 < 0.01
    (:218-254) and exact agreement with the f64 evaluation of the fixture graph."""
    ctx = b.Context(0)
    for week in (22.0, 23.0, 24.0):
        x = np.array([[60.1699, 24.9384, week]], np.float32)
        got = run(ctx, x, W_FIX, B_FIX, b.ACT_SIGMOID)[0]
        ref = 1.0 / (1.0 + np.exp(-(x.astype(np.float64) @ W_FIX.astype(np.float64) + B_FIX)))[0]
        assert np.abs(got - ref).max() <= 1e-6
        assert got[3] < 0.01 and got[0] > 0.5
    ctx.close()


@pytest.mark.parametrize("B,K,N,act", [(64, 1024, 38, b.ACT_SOFTMAX), (7, 1024, 11, b.ACT_SIGMOID), (1, 3, 12012, b.ACT_SIGMOID),
                                       (5, 33, 17, b.ACT_NONE)])
def test_dense_shapes(B, K, N, act):
    rng = np.random.default_rng(B + N)
    x = rng.standard_normal((B, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ctx = b.Context(0)
    got = run(ctx, x, W, bias, act)
    z = x.astype(np.float64) @ W.astype(np.float64) + bias
    if act == b.ACT_SIGMOID:
        ref = 1 / (1 + np.exp(-z))
    elif act == b.ACT_SOFTMAX:
        e = np.exp(z - z.max(axis=1, keepdims=True)); ref = e / e.sum(axis=1, keepdims=True)
    else:
        ref = z
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    ctx.close()


def test_per_class_calibration_then_post():
    """bb_calibrate_run: out = a[c] * logit + b[c] per class (the standard Platt form; the reference's BSG calibration is
    inside birdnet-onnx and unpinned), in place, then the usual sigmoid post step on the calibrated logits."""
    import ctypes as C

    import torch
    from birda_b200 import _lib
    from oracle import post as opost
    ctx = b.Context(0)
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((40, 265)) * 2 - 4).astype(np.float32)
    a = (0.5 + rng.random(265)).astype(np.float32); bb = (rng.standard_normal(265) * 0.5).astype(np.float32)
    d = torch.from_numpy(x).cuda(); da = torch.from_numpy(a).cuda(); db = torch.from_numpy(bb).cuda()
    _lib.check(_lib.lib.bb_calibrate_run(ctx.handle, C.c_void_p(d.data_ptr()), 40, 265, C.c_void_p(da.data_ptr()), C.c_void_p(db.data_ptr()),
                                         C.c_void_p(d.data_ptr())), ctx.handle)
    ctx.sync()
    want = np.float32(a) * x + bb                       # fmaf vs mul+add: one rounding apart
    got = d.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    idx, conf, cnt = ctx.post_run(d.data_ptr(), 40, 265, 40, b.PostConfig(min_confidence=0.1))
    ref = opost.post_process(got, 40, opost.ACT_SIGMOID, 0.1, 5)
    for r in range(40):
        assert [int(i) for i in idx[r, : cnt[r]]] == [i for i, _ in ref[r]]
    ctx.close()
