"""The C restatement (own f32 FFT) against the numpy restatement: two independent oracles agree."""
import numpy as np
import pytest

from birda_b200.synth import synth_logits, synth_pcm
from oracle import cport
from oracle import frontend as ofe
from oracle import post as opost


@pytest.mark.parametrize("sr,tr", [(44_100, 48_000), (48_000, 32_000), (44_100, 32_000), (96_000, 48_000), (16_000, 48_000)])
def test_c_resampler_matches_numpy(sr, tr):
    p = ofe.make_plan(sr, tr)
    n_in, n_out, cutoff, taps = cport.resampler_info(sr, tr)
    assert (n_in, n_out) == (p.n_in, p.n_out) and abs(float(cutoff) - float(p.cutoff)) < 2e-7
    assert np.abs(taps - p.taps).max() <= 4e-7 * np.abs(p.taps).max()
    x = (np.random.default_rng(sr).standard_normal(n_in * 5 + 333) * 0.2).astype(np.float32)
    a = cport.resample(x, sr, tr)
    b64 = ofe.resample(x, sr, tr, precision="f64")
    assert a.size == b64.size
    scale = np.sqrt(np.mean(b64.astype(np.float64) ** 2))
    assert np.abs(a - b64).max() <= 5e-6 * scale


def test_c_frontend_matches_numpy_tables_and_samples():
    pcm = synth_pcm(2, 9.3, 44_100, 2)
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 72_000, precision="f64")
    seg, ss, st, et = cport.frontend(pcm, 2, 44_100, 48_000, 144_000, 72_000, threads=4)
    assert np.array_equal(ss, ref.start_sample)
    assert st.tobytes() == ref.start_time.tobytes() and et.tobytes() == ref.end_time.tobytes()
    rms = np.sqrt(np.mean(ref.segments.astype(np.float64) ** 2, axis=1, keepdims=True))
    err = (np.abs(seg - ref.segments) / np.maximum(np.abs(ref.segments), np.maximum(rms, 1e-30))).max()
    assert err <= 1e-5
    # no-resample path is bit exact
    pcm = synth_pcm(3, 8.0, 48_000, 3, np.int32)
    ref = ofe.decode_and_stream(pcm, 3, 48_000, 48_000, 144_000, 48_000)
    seg, ss, st, et = cport.frontend(pcm, 3, 48_000, 48_000, 144_000, 48_000)
    assert np.array_equal(seg.view(np.uint32), ref.segments.view(np.uint32)) and np.array_equal(ss, ref.start_sample)


@pytest.mark.parametrize("keep_unmatched,rerank", [(True, False), (False, False), (True, True)])
def test_c_post_matches_numpy(keep_unmatched, rerank):
    C = 6522
    rng = np.random.default_rng(9)
    x = synth_logits(9, 48, C)
    mask = (rng.random(C) ** 2).astype(np.float32); mask[rng.choice(C, 305, replace=False)] = np.nan
    idx, conf, cnt = cport.post(x, 48, 1, 0.1, 5, mask, None, 0.01, keep_unmatched, rerank, threads=2)
    ref = opost.post_process(x, 48, opost.ACT_SIGMOID, 0.1, 5, mask, opost.FilterSettings(0.01, keep_unmatched, rerank))
    for r in range(48):
        if r == 2:
            continue        # value planted exactly at logit(min_conf): a threshold-boundary tie (f32 vs f64 sigmoid)
        assert [int(i) for i in idx[r, : cnt[r]]] == [i for i, _ in ref[r]]
        assert all(abs(float(conf[r, j]) - float(ref[r][j][1])) <= 1e-6 for j in range(int(cnt[r])))
