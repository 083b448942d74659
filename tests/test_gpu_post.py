"""Parity of the CUDA post-inference step (through the C ABI) against the oracle."""
import numpy as np
import pytest

import birda_b200 as b
from birda_b200.synth import synth_logits
from oracle import post as opost

pytestmark = pytest.mark.gpu
CONF_TOL = 1e-4    # north_star: confidences within 1e-4 absolute


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0)
    yield c
    c.close()


def gpu_post(ctx, scores, valid, cfg, mask=None, keep=None):
    import torch
    d = torch.from_numpy(scores).cuda()
    dm = torch.from_numpy(mask).cuda() if mask is not None else None
    dk = torch.from_numpy(keep.astype(np.uint8)).cuda() if keep is not None else None
    idx, conf, cnt = ctx.post_run(d.data_ptr(), scores.shape[0], scores.shape[1], valid, cfg,
                                  dm.data_ptr() if dm is not None else None, dk.data_ptr() if dk is not None else None)
    return [[(int(idx[r, j]), np.float32(conf[r, j])) for j in range(int(cnt[r]))] for r in range(valid)]


def compare(got, ref, conf_ref_all, min_conf):
    """Identical detection sets except rows whose decision sits within CONF_TOL of a boundary."""
    assert len(got) == len(ref)
    for r, (g, o) in enumerate(zip(got, ref)):
        gi, oi = [i for i, _ in g], [i for i, _ in o]
        if gi != oi:
            # tolerated only at a boundary: some confidence within tol of min_conf or of its neighbour
            c = np.sort(conf_ref_all[r])[::-1][:8].astype(np.float64)
            near_thr = np.any(np.abs(c - min_conf) <= CONF_TOL)
            near_tie = np.any(np.abs(np.diff(c)) <= CONF_TOL)
            assert near_thr or near_tie, (r, g, o)
            continue
        for (i1, c1), (i2, c2) in zip(g, o):
            assert abs(float(c1) - float(c2)) <= CONF_TOL, (r, i1, c1, c2)


@pytest.mark.parametrize("classes,act", [(6522, b.ACT_SIGMOID), (14795, b.ACT_SOFTMAX), (11560, b.ACT_NONE), (265, b.ACT_SIGMOID)])
def test_post_plain(ctx, classes, act):
    x = synth_logits(classes, 70, classes)
    if act == b.ACT_NONE:
        x = opost.activate(x, opost.ACT_SIGMOID)            # graph already applied the sigmoid
    if act == b.ACT_SOFTMAX:
        x[6:, :] *= 3.0                                     # peaky rows so something clears 0.1
    cfg = b.PostConfig(activation=act, min_confidence=0.1)
    got = gpu_post(ctx, x, 64, cfg)                         # rows 64..69 are batch padding
    ref = opost.post_process(x, 64, act, 0.1, 5)
    compare(got, ref, opost.activate(x, act), 0.1)
    assert sum(len(r) for r in ref) > 50


@pytest.mark.parametrize("classes,act", [(6522, b.ACT_SIGMOID), (14795, b.ACT_SOFTMAX), (1001, b.ACT_SIGMOID)])
def test_post_many_rows(ctx, classes, act):
    """>= 1024 rows take the 64-thread-CTA variant of K3 (one launch over a whole file's windows); the row stride of
    C = 6522 / 14795 / 1001 floats also walks through every 16-byte misalignment of the row start."""
    rows = 1200
    x = synth_logits(classes + 1, rows, classes)
    if act == b.ACT_SOFTMAX:
        x[6:, :] *= 3.0
    cfg = b.PostConfig(activation=act, min_confidence=0.1)
    got = gpu_post(ctx, x, rows, cfg)
    ref = opost.post_process(x, rows, act, 0.1, 5)
    compare(got, ref, opost.activate(x, act), 0.1)
    assert sum(len(r) for r in ref) > 500


def test_post_adversarial_rows_exact(ctx):
    x = synth_logits(1, 8, 6522)
    got = gpu_post(ctx, x, 8, b.PostConfig(min_confidence=0.1))
    ref = opost.post_process(x, 8, opost.ACT_SIGMOID, 0.1, 5)
    assert got[0] == [] and ref[0] == []
    assert [i for i, _ in got[1]] == [0, 5, 17, 300, 301]          # exact ties: lower index wins
    assert [i for i, _ in got[4]] == [0, 1, 2, 3, 4]
    assert [i for i, _ in got[5]] == [6519, 6520, 6521]
    assert [i for i, _ in got[3]] == [i for i, _ in ref[3]] and len(got[3]) == 5


@pytest.mark.parametrize("keep_unmatched,rerank", [(True, False), (False, False), (True, True), (False, True)])
def test_post_range_mask(ctx, keep_unmatched, rerank):
    """C5's filter: synthetic mask U(0,1)^2 with 305 NaN entries, threshold 0.01."""
    C = 6522
    rng = np.random.default_rng(44)
    x = synth_logits(5, 96, C)
    mask = (rng.random(C) ** 2).astype(np.float32)
    mask[rng.choice(C, 305, replace=False)] = np.nan
    mask[rng.choice(C, 400, replace=False)] = 0.0
    mask[7] = np.float32(0.01)                              # exactly at the (inclusive) threshold
    x[10, 7] = 4.0
    cfg = b.PostConfig(min_confidence=0.1, range_threshold=0.01, keep_unmatched=keep_unmatched, rerank=rerank)
    got = gpu_post(ctx, x, 96, cfg, mask=mask)
    ref = opost.post_process(x, 96, opost.ACT_SIGMOID, 0.1, 5, mask,
                             opost.FilterSettings(0.01, keep_unmatched, rerank))
    conf_all = opost.activate(x, opost.ACT_SIGMOID)
    if rerank:
        conf_all = conf_all * np.nan_to_num(mask, nan=0.0)[None, :]
    compare(got, ref, conf_all, 0.1)
    if not rerank:
        assert any(7 in [i for i, _ in r] for r in got)    # score exactly at the inclusive threshold survives
    plain = opost.post_process(x, 96, opost.ACT_SIGMOID, 0.1, 5)
    assert sum(map(len, ref)) < sum(map(len, plain))        # the mask removed something


def test_post_species_list(ctx):
    C = 6522
    x = synth_logits(6, 40, C)
    keep = np.zeros(C, bool); keep[::3] = True
    got = gpu_post(ctx, x, 40, b.PostConfig(min_confidence=0.1), keep=keep)
    ref = opost.post_process(x, 40, opost.ACT_SIGMOID, 0.1, 5, species_keep=keep)
    compare(got, ref, opost.activate(x, opost.ACT_SIGMOID), 0.1)
    assert all(i % 3 == 0 for r in got for i, _ in r)


def test_post_min_conf_zero_and_topk(ctx):
    x = synth_logits(8, 16, 1000)
    for k in (1, 3, 8):
        got = gpu_post(ctx, x, 16, b.PostConfig(min_confidence=0.0, top_k=k))
        ref = opost.post_process(x, 16, opost.ACT_SIGMOID, 0.0, k)
        compare(got, ref, opost.activate(x, opost.ACT_SIGMOID), 0.0)
        assert all(len(r) == k for r in got)


def test_post_errors(ctx):
    import torch
    d = torch.zeros(4, 10, device="cuda")
    with pytest.raises(b.BirdaError):
        ctx.post_run(d.data_ptr(), 4, 10, 5, b.PostConfig())          # valid > B
    with pytest.raises(b.BirdaError):
        ctx.post_run(d.data_ptr(), 4, 10, 4, b.PostConfig(top_k=9))
    with pytest.raises(b.BirdaError):
        ctx.post_run(d.data_ptr(), 4, 10, 4, b.PostConfig(), d.data_ptr(), d.data_ptr())


def _plateau_rows(C, seed):
    """Rows where the order is decided on a plateau of the activation: saturated sigmoids (every score above ~16.7 is
    exactly 1.0f, the class index decides), scores around the 8.0 knee of the kernel's threshold rule, confidences that
    underflow to 0 (index order again), and a row with fewer non-trivial scores than top_k."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((12, C)) * 2.0 - 6.0).astype(np.float32)
    x[0, :] = -20.0; x[0, rng.choice(C, 40, replace=False)] = rng.uniform(18.0, 30.0, 40).astype(np.float32)   # all exactly 1.0f
    x[1, :] = -20.0; x[1, rng.choice(C, 40, replace=False)] = rng.uniform(7.9, 8.2, 40).astype(np.float32)
    x[2, :] = -100.0; x[2, rng.choice(C, 3, replace=False)] = 2.0
    x[3, :] = -200.0
    x[4, :] = -20.0; x[4, C - 1] = 25.0; x[4, 1] = 20.0; x[4, 0] = 12.0          # 1.0f, 1.0f, clearly below
    x[5, rng.choice(C, 300, replace=False)] = rng.uniform(18.0, 40.0, 300).astype(np.float32)   # 300 ties at 1.0f
    x[6, :] = np.float32(np.log(0.1 / 0.9))                                      # the whole row at the threshold
    x[7, :] = -np.inf; x[7, 5] = 0.0
    x[8, ::2] = np.nan                                                           # NaN scores never win
    return x


@pytest.mark.parametrize("rows", [12, 1100])
@pytest.mark.parametrize("min_conf,top_k", [(0.1, 5), (0.0, 5), (0.0, 8), (0.5, 1)])
def test_post_activation_plateaus_exact(ctx, rows, min_conf, top_k):
    """K3 filters a row with a threshold derived from the row itself (csrc/k3_post.cu: lowered()); wherever confidences tie
    the winner is the lower class index, whatever the scores behind the tie were.  Index lists must match exactly."""
    C = 6522
    x = _plateau_rows(C, 3)
    if rows > 12:                                           # the 64-thread CTA variant: threshold from a third of the row
        x = np.concatenate([x, synth_logits(9, rows - 12, C)], axis=0)
    got = gpu_post(ctx, x, rows, b.PostConfig(min_confidence=min_conf, top_k=top_k))
    ref = opost.post_process(x, rows, opost.ACT_SIGMOID, min_conf, top_k)
    for r in (0, 2, 3, 4, 5, 7):                            # ties that are exact by construction
        assert [i for i, _ in got[r]] == [i for i, _ in ref[r]], (r, got[r], ref[r])
    assert [i for i, _ in got[4]][:2] == [1, C - 1][:top_k]
    compare(got, ref, np.nan_to_num(opost.activate(x, opost.ACT_SIGMOID), nan=-1.0), min_conf)


@pytest.mark.parametrize("rows", [10, 1100])
def test_post_softmax_underflow_and_ties(ctx, rows):
    C = 1001
    rng = np.random.default_rng(12)
    x = rng.standard_normal((rows, C)).astype(np.float32)
    x[0, 500] = 200.0                                       # everything else underflows to exactly 0: index order after the peak
    x[1, :] = 3.0                                           # uniform row: 1/C everywhere
    x[2, 7] = 200.0; x[2, 9] = 200.0                        # two equal peaks, the rest exactly 0
    x[3, :] = -np.inf; x[3, 4] = 1.0
    for min_conf, k in ((0.0, 5), (0.0005, 5), (0.3, 3)):
        got = gpu_post(ctx, x, rows, b.PostConfig(activation=b.ACT_SOFTMAX, min_confidence=min_conf, top_k=k))
        ref = opost.post_process(x, rows, opost.ACT_SOFTMAX, min_conf, k)
        for r in range(4):
            assert [i for i, _ in got[r]] == [i for i, _ in ref[r]], (min_conf, r, got[r], ref[r])
        compare(got, ref, opost.activate(x, opost.ACT_SOFTMAX), min_conf)


def test_post_identity_activation_negative_scores(ctx):
    """ACT_NONE with scores and a threshold below zero (graph outputs that are not probabilities)."""
    C = 777
    rng = np.random.default_rng(13)
    x = (rng.standard_normal((1100, C)) - 3.0).astype(np.float32)
    x[0, :] = -7.0                                          # all tied: lowest indices
    got = gpu_post(ctx, x, 1100, b.PostConfig(activation=b.ACT_NONE, min_confidence=-2.5, top_k=4))
    ref = opost.post_process(x, 1100, opost.ACT_NONE, -2.5, 4)
    assert got[0] == [] and ref[0] == []
    compare(got, ref, x, -2.5)
    got = gpu_post(ctx, x, 1100, b.PostConfig(activation=b.ACT_NONE, min_confidence=-10.0, top_k=4))
    ref = opost.post_process(x, 1100, opost.ACT_NONE, -10.0, 4)
    assert [i for i, _ in got[0]] == [0, 1, 2, 3]
    compare(got, ref, x, -10.0)


@pytest.mark.parametrize("classes,act", [(6522, b.ACT_SIGMOID), (14795, b.ACT_SOFTMAX)])
def test_post_few_hundred_rows(ctx, classes, act):
    """128..1023 rows take the 128-thread-CTA variant of K3."""
    rows = 300
    x = synth_logits(classes + 2, rows, classes)
    if act == b.ACT_SOFTMAX:
        x[6:, :] *= 3.0
    got = gpu_post(ctx, x, rows, b.PostConfig(activation=act, min_confidence=0.1))
    ref = opost.post_process(x, rows, act, 0.1, 5)
    compare(got, ref, opost.activate(x, act), 0.1)
    assert sum(len(r) for r in ref) > 100
