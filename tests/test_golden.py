"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle on seeded inputs).
CPU: the oracle still reproduces them.  GPU: the CUDA path through the C ABI matches them."""
import glob
import os

import numpy as np
import pytest

import birda_b200 as b
from oracle import frontend as ofe
from oracle import melspec as om
from oracle import post as opost
from oracle import rules as orules

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRONTEND = sorted(glob.glob(os.path.join(GOLD, "frontend_*.npz")))


def rel_err(got, ref):
    ref = ref.astype(np.float64)
    rms = np.sqrt(np.mean(ref ** 2, axis=1, keepdims=True))
    return float((np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref), np.maximum(rms, 1e-30))).max())


def test_fixture_set_is_complete():
    assert len(FRONTEND) == 5
    for f in ("post_sigmoid_265.npz", "segment_tables.npz", "melspec_512_32.npz"):
        assert os.path.exists(os.path.join(GOLD, f))


@pytest.mark.parametrize("path", FRONTEND, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_frontend_golden(path):
    g = np.load(path)
    sr, ch, tr, seg, ovl = (int(v) for v in g["params"])
    for precision in ("f64", "f32"):
        r = ofe.decode_and_stream(g["pcm"], ch, sr, tr, seg, ovl, precision=precision)
        assert np.array_equal(r.start_sample, g["expected_start_sample"])
        assert r.start_time.tobytes() == g["expected_start_time"].tobytes()
        assert r.end_time.tobytes() == g["expected_end_time"].tobytes()
        if sr == tr:
            assert np.array_equal(r.segments, g["expected_segments"])
        else:
            assert rel_err(r.segments, g["expected_segments"]) <= 1e-5


def test_oracle_reproduces_post_tables_melspec_golden():
    g = np.load(os.path.join(GOLD, "post_sigmoid_265.npz"))
    for tag, kw in {"plain": {}, "mask_keep": dict(mask=g["mask"], settings=opost.FilterSettings(0.01, True, False)),
                    "mask_rerank": dict(mask=g["mask"], settings=opost.FilterSettings(0.01, False, True))}.items():
        rows = opost.post_process(g["scores"], 24, opost.ACT_SIGMOID, 0.1, 5, kw.get("mask"), kw.get("settings"))
        for r, row in enumerate(rows):
            assert len(row) == g[f"expected_{tag}_count"][r]
            assert [i for i, _ in row] == list(g[f"expected_{tag}_index"][r, : len(row)])
            assert np.array_equal(np.array([c for _, c in row], np.float32), g[f"expected_{tag}_conf"][r, : len(row)])
    t = np.load(os.path.join(GOLD, "segment_tables.npz"))
    for key in t.files:
        total, seg, ovl = (int(v) for v in key.split("_")[1:])
        got = np.array([(w.start_sample, w.take) for w in orules.next_segment_table(total, seg, ovl)], np.int64).reshape(-1, 2)
        assert np.array_equal(got, t[key]), key
    m = np.load(os.path.join(GOLD, "melspec_512_32.npz"))
    n_fft, hop, n_frames = (int(v) for v in m["params"])
    ref = om.melspec(m["segments"], n_fft, hop, n_frames, m["window"], m["mel_weights"])
    assert np.allclose(ref, m["expected"], rtol=1e-6, atol=1e-9 * np.abs(ref).max())


def test_library_rules_reproduce_segment_tables_golden():
    """The product's host rules (C++ in the library, no GPU needed) against the frozen tables."""
    t = np.load(os.path.join(GOLD, "segment_tables.npz"))
    for key in t.files:
        total, seg, ovl = (int(v) for v in key.split("_")[1:])
        assert b.rules.segment_count(total, seg, ovl) == t[key].shape[0], key
        if t[key].shape[0]:
            start, take = b.rules.segment_table(total, seg, ovl)[:2]
            assert np.array_equal(np.asarray(start, np.int64), t[key][:, 0]) and np.array_equal(np.asarray(take, np.int64), t[key][:, 1]), key


# ----------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FRONTEND, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_frontend_matches_golden(ctx, path):
    g = np.load(path)
    sr, ch, tr, seg, ovl = (int(v) for v in g["params"])
    plan = b.FrontEndPlan(ctx, sr, ch, b.FMT_S16, tr, seg, ovl)
    res = plan.run(g["pcm"], pad_to_batch=4)
    ctx.sync()
    out = res.torch().cpu().numpy().copy()
    plan.close()
    assert res.nseg == g["expected_segments"].shape[0] and np.array_equal(res.start_sample, g["expected_start_sample"])
    assert res.start_time.tobytes() == g["expected_start_time"].tobytes() and res.end_time.tobytes() == g["expected_end_time"].tobytes()
    assert not out[res.nseg:].any()
    if sr == tr:
        assert np.array_equal(out[: res.nseg], g["expected_segments"])          # bit exact
    else:
        assert rel_err(out[: res.nseg], g["expected_segments"]) <= 1e-5         # north_star tolerance


@pytest.mark.gpu
def test_cuda_post_matches_golden(ctx):
    import torch
    g = np.load(os.path.join(GOLD, "post_sigmoid_265.npz"))
    d = torch.from_numpy(g["scores"]).cuda()
    dm = torch.from_numpy(g["mask"]).cuda()
    conf_all = opost.activate(g["scores"], opost.ACT_SIGMOID)
    for tag, cfg, mask in (("plain", b.PostConfig(min_confidence=0.1), None),
                           ("mask_keep", b.PostConfig(min_confidence=0.1, range_threshold=0.01, keep_unmatched=True, rerank=False), dm),
                           ("mask_rerank", b.PostConfig(min_confidence=0.1, range_threshold=0.01, keep_unmatched=False, rerank=True), dm)):
        idx, conf, cnt = ctx.post_run(d.data_ptr(), 24, 265, 24, cfg, mask.data_ptr() if mask is not None else None)
        for r in range(24):
            n = int(g[f"expected_{tag}_count"][r])
            ei, ec = g[f"expected_{tag}_index"][r, :n], g[f"expected_{tag}_conf"][r, :n]
            top = np.sort(conf_all[r])[::-1][:8].astype(np.float64)               # boundary rows: a confidence within tol of min_conf or of its neighbour
            near = np.any(np.abs(top - 0.1) <= 1e-4) or np.any(np.abs(np.diff(top)) <= 1e-4)
            if int(cnt[r]) != n or list(map(int, idx[r, :n])) != list(map(int, ei)):
                assert near, (tag, r)
                continue
            assert np.all(np.abs(conf[r, :n] - ec) <= 1e-4), (tag, r)


@pytest.mark.gpu
def test_cuda_melspec_matches_golden(ctx):
    import torch
    m = np.load(os.path.join(GOLD, "melspec_512_32.npz"))
    n_fft, hop, n_frames = (int(v) for v in m["params"])
    d = torch.from_numpy(m["segments"]).cuda()
    out = torch.empty((2, 32, n_frames), device="cuda")
    ms = b.MelSpec(ctx, n_fft, hop, n_frames, m["window"], m["mel_weights"])
    ms.run(d.data_ptr(), 2, m["segments"].shape[1], out.data_ptr())
    ctx.sync()
    ms.close()
    ref = m["expected"].astype(np.float64)
    assert (np.abs(out.cpu().numpy() - ref) / np.abs(ref).max(axis=(1, 2), keepdims=True)).max() <= 2e-5
