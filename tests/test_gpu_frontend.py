"""Parity of the CUDA front end (through the C ABI) against the oracle on a real GPU."""
import numpy as np
import pytest

import birda_b200 as b
from birda_b200.synth import synth_pcm
from oracle import frontend as ofe
from oracle import rules as orules

pytestmark = pytest.mark.gpu

RESAMPLE_TOL = 1e-5     # north_star: resampled samples within 1e-5 relative


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0)
    yield c
    c.close()


def run_gpu(ctx, pcm, channels, sr, tr, seg, ovl, fmt, pad=0, device_input=False):
    plan = b.FrontEndPlan(ctx, sr, channels, fmt, tr, seg, ovl)
    if device_input:
        import torch
        t = torch.from_numpy(pcm).cuda()
        res = plan.run(t, pad_to_batch=pad)
    else:
        res = plan.run(pcm, pad_to_batch=pad)
    ctx.sync()
    out = res.torch().cpu().numpy().copy()
    plan.close()
    return res, out


def assert_tables(res, ref):
    assert res.nseg == ref.segments.shape[0]
    assert np.array_equal(res.start_sample, ref.start_sample)
    assert res.start_time.tobytes() == ref.start_time.tobytes()
    assert res.end_time.tobytes() == ref.end_time.tobytes()


def rel_err(got, ref):
    """max |got-ref| / max(|ref|, rms(ref)) per SURVEY §8d parity gates (row-wise rms)."""
    rms = np.sqrt(np.mean(ref.astype(np.float64) ** 2, axis=1, keepdims=True))
    scale = np.maximum(np.abs(ref), np.maximum(rms, 1e-30))
    return float((np.abs(got.astype(np.float64) - ref) / scale).max())


# ------------------------------------------------------------------ K1: no resampling, bit exact
@pytest.mark.parametrize("dtype,fmt", [(np.int16, b.FMT_S16), (np.int32, b.FMT_S32), (np.float32, b.FMT_F32)])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_pack_bit_exact_formats(ctx, dtype, fmt, channels):
    pcm = synth_pcm(7 + channels, 10.7, 48_000, channels, dtype)
    ref = ofe.decode_and_stream(pcm, channels, 48_000, 48_000, 144_000, 48_000)
    res, out = run_gpu(ctx, pcm, channels, 48_000, 48_000, 144_000, 48_000, fmt)
    assert_tables(res, ref)
    assert np.array_equal(out[: res.nseg].view(np.uint32), ref.segments.view(np.uint32))


def test_c1_birdnet_60s_mono(ctx):
    """BASELINE config 1: 60 s 48 kHz mono, overlap 0 -> 20 segments, batch 8 padding."""
    pcm = synth_pcm(1, 60.0, 48_000, 1)
    ref = ofe.decode_and_stream(pcm, 1, 48_000, 48_000, 144_000, 0)
    res, out = run_gpu(ctx, pcm, 1, 48_000, 48_000, 144_000, 0, b.FMT_S16, pad=8)
    assert res.nseg == 20 and res.rows == 24
    assert_tables(res, ref)
    assert np.array_equal(out[:20], ref.segments)
    assert not out[20:].any()                      # padding rows are silence (processor.rs:239-260)


def test_c4_bat_256k(ctx):
    """BASELINE config 4 (short): 256 kHz mono, 144000-sample windows, 36000 overlap, no resample."""
    pcm = synth_pcm(4, 5.3, 256_000, 1, bat=True)
    seg, ovl = b.rules.segment_samples(0.5625, 0.0, 256_000, bat_mode=True)
    ref = ofe.decode_and_stream(pcm, 1, 256_000, 256_000, seg, ovl)
    res, out = run_gpu(ctx, pcm, 1, 256_000, 256_000, seg, ovl, b.FMT_S16, device_input=True)
    assert_tables(res, ref)
    assert np.array_equal(out[: res.nseg], ref.segments)


@pytest.mark.parametrize("total,seg,ovl", [(1000, 400, 100), (1050, 400, 100), (50, 400, 100), (400, 400, 0),
                                           (401, 400, 399), (12_345, 1001, 333), (7, 3, 2), (0, 100, 10)])
def test_tail_semantics_and_unaligned(ctx, total, seg, ovl):
    """take <= overlap tails, odd sizes (scalar path), empty input."""
    rng = np.random.default_rng(total)
    pcm = rng.integers(-32768, 32767, total * 2, dtype=np.int16)
    ref = ofe.decode_and_stream(pcm, 2, 48_000, 48_000, seg, ovl)
    res, out = run_gpu(ctx, pcm, 2, 48_000, 48_000, seg, ovl, b.FMT_S16) if total else (None, None)
    if total == 0:
        plan = b.FrontEndPlan(ctx, 48_000, 2, b.FMT_S16, 48_000, seg, ovl)
        r = plan.run(pcm)
        assert r.nseg == 0 and r.rows == 0
        return
    assert_tables(res, ref)
    assert np.array_equal(out[: res.nseg], ref.segments)


def test_negative_zero_and_extremes_f32(ctx):
    x = np.array([-0.0, 0.0, -0.0, -0.0, 1.0, -1.0, 3.4e38, 3.4e38, 1e-45, -1e-45, np.inf, -np.inf] * 50, np.float32)
    for ch in (1, 2, 3):
        pcm = x[: (x.size // ch) * ch]
        with np.errstate(all="ignore"):
            ref = ofe.decode_and_stream(pcm, ch, 48_000, 48_000, 64, 16)
        res, out = run_gpu(ctx, pcm, ch, 48_000, 48_000, 64, 16, b.FMT_F32)
        got = out[: res.nseg]
        nan = np.isnan(ref.segments)                       # inf + -inf: NaN payload/sign is not part of parity
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got.view(np.uint32)[~nan], ref.segments.view(np.uint32)[~nan])


def test_streaming_pieces_equal_whole_file(ctx):
    """is_eof=0 pieces + consumed_frames reproduce the whole-file result."""
    pcm = synth_pcm(9, 20.0, 48_000, 1)
    seg, ovl = 144_000, 72_000
    ref = ofe.decode_and_stream(pcm, 1, 48_000, 48_000, seg, ovl)
    plan = b.FrontEndPlan(ctx, 48_000, 1, b.FMT_S16, 48_000, seg, ovl)
    rows, ss, pos = [], [], 0
    piece = 48_000 * 7
    while True:
        end = min(pcm.size, pos + piece)
        eof = end == pcm.size
        r = plan.run(pcm[pos:end], first_start_sample=pos, is_eof=eof)
        ctx.sync()
        if r.nseg:
            rows.append(r.torch().cpu().numpy()[: r.nseg].copy()); ss.append(r.start_sample.copy())
        if eof:
            break
        pos += r.consumed_frames
        assert r.consumed_frames > 0
    assert np.array_equal(np.concatenate(ss), ref.start_sample)
    assert np.array_equal(np.concatenate(rows), ref.segments)


# ------------------------------------------------------------------ K2: resampled, 1e-5 relative
@pytest.mark.parametrize("sr,tr,seg,ovl,channels,seconds", [
    (44_100, 48_000, 144_000, 72_000, 2, 9.3),     # C2 shape (K2's own blocking, s16 stereo staging)
    (44_100, 48_000, 144_000, 36_001, 1, 8.7),     # the same plan from mono s16 (word-pair staging, odd block starts)
    (48_000, 32_000, 160_000, 0, 1, 12.1),         # C3 shape (radix 19)
    (22_050, 48_000, 144_000, 0, 1, 7.0),
    (96_000, 48_000, 144_000, 48_000, 2, 7.5),
    (16_000, 48_000, 144_000, 0, 1, 6.5),
    (32_000, 48_000, 144_000, 62_399, 1, 6.5),
    (44_100, 32_000, 160_000, 0, 2, 11.0),
])
def test_resample_parity(ctx, sr, tr, seg, ovl, channels, seconds):
    pcm = synth_pcm(sr % 97, seconds, sr, channels)
    ref = ofe.decode_and_stream(pcm, channels, sr, tr, seg, ovl, precision="f64")
    res, out = run_gpu(ctx, pcm, channels, sr, tr, seg, ovl, b.FMT_S16, pad=4)
    assert_tables(res, ref)
    assert res.rows % 4 == 0 and not out[res.nseg:].any()
    err = rel_err(out[: res.nseg], ref.segments.astype(np.float64))
    assert err <= RESAMPLE_TOL, err
    # and against the f32-faithful oracle
    ref32 = ofe.decode_and_stream(pcm, channels, sr, tr, seg, ovl, precision="f32")
    assert rel_err(out[: res.nseg], ref32.segments.astype(np.float64)) <= RESAMPLE_TOL


def test_resample_generic_kernel_path(ctx, monkeypatch):
    """Rates without a compile-time plan use the generic Stockham kernel; force it for a planned pair too."""
    monkeypatch.setenv("BIRDA_K2_GENERIC", "1")
    pcm = synth_pcm(5, 8.0, 44_100, 2)
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 72_000, precision="f64")
    res, out = run_gpu(ctx, pcm, 2, 44_100, 48_000, 144_000, 72_000, b.FMT_S16)
    assert_tables(res, ref)
    assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL
    monkeypatch.delenv("BIRDA_K2_GENERIC")
    # 24 kHz -> 48 kHz (1024 -> 2048 blocks) has no compile-time plan
    pcm = synth_pcm(6, 7.0, 24_000, 1)
    ref = ofe.decode_and_stream(pcm, 1, 24_000, 48_000, 144_000, 0, precision="f64")
    res, out = run_gpu(ctx, pcm, 1, 24_000, 48_000, 144_000, 0, b.FMT_S16)
    assert_tables(res, ref)
    assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL


@pytest.mark.parametrize("sr,channels", [(12_345, 1), (8_001, 2), (37_000, 1)])
def test_resample_rates_with_large_prime_blocks(ctx, sr, channels):
    """Rates whose blocks have a prime factor > 31 (12 345 Hz: 2 * 823; 8 001 Hz: 3 * 7 * 127; 37 kHz: 37 * 28): the
    reference resamples whatever rubato accepts (src/audio/resample.rs:19-28); here the fallback kernel evaluates the
    big prime's stage as a direct DFT.  Same gate as every other pair."""
    pcm = synth_pcm(70 + channels, 7.3, sr, channels)
    ref = ofe.decode_and_stream(pcm, channels, sr, 48_000, 144_000, 0, precision="f64")
    res, out = run_gpu(ctx, pcm, channels, sr, 48_000, 144_000, 0, b.FMT_S16)
    assert_tables(res, ref)
    assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL


@pytest.mark.parametrize("dtype,fmt,channels", [(np.int32, b.FMT_S32, 2), (np.float32, b.FMT_F32, 1), (np.int16, b.FMT_S16, 3)])
def test_resample_other_formats(ctx, dtype, fmt, channels):
    pcm = synth_pcm(40 + channels, 6.4, 44_100, channels, dtype)
    ref = ofe.decode_and_stream(pcm, channels, 44_100, 48_000, 144_000, 0, precision="f64")
    res, out = run_gpu(ctx, pcm, channels, 44_100, 48_000, 144_000, 0, fmt)
    assert_tables(res, ref)
    assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL


@pytest.mark.parametrize("sr,tr,seg,channels,skip", [
    (44_100, 48_000, 144_000, 2, 1),      # stereo s16, device pointer 4 bytes past a 16-byte boundary: every block's staging is 4-byte aligned only
    (44_100, 48_000, 144_000, 1, 1),      # mono s16 starting mid-word (funnel-shift path)
    (44_100, 48_000, 144_000, 1, 2),
    (48_000, 32_000, 160_000, 1, 3),
    (96_000, 48_000, 144_000, 2, 3),      # padded power-of-two plan with an odd frame offset
    (48_000, 48_000, 144_000, 1, 1),      # K1: row starts not 16-byte aligned -> general path
    (48_000, 48_000, 144_000, 2, 3),
])
def test_device_pointer_at_odd_offsets(ctx, sr, tr, seg, channels, skip):
    """The ABI takes any device pointer: interior fast paths, cp.async staging and vector loads must not assume alignment."""
    import torch
    full = synth_pcm(70 + skip, 9.0, sr, channels)
    t = torch.from_numpy(full).cuda()
    pcm = full[skip * channels:]
    frames = pcm.size // channels
    ref = ofe.decode_and_stream(pcm, channels, sr, tr, seg, 0, precision="f64")
    plan = b.FrontEndPlan(ctx, sr, channels, b.FMT_S16, tr, seg, 0)
    res = plan.run(t.data_ptr() + skip * channels * 2, frames, is_device=True)
    ctx.sync()
    out = res.torch().cpu().numpy().copy()
    plan.close()
    assert_tables(res, ref)
    if sr == tr:
        assert np.array_equal(out[: res.nseg], ref.segments)
    else:
        assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL


def test_resample_many_windows_split_runs(ctx):
    """Few windows -> the kernel splits each window into runs of blocks (recomputed carry)."""
    pcm = synth_pcm(21, 3.4, 44_100, 1)
    ref = ofe.decode_and_stream(pcm, 1, 44_100, 48_000, 144_000, 0, precision="f64")
    res, out = run_gpu(ctx, pcm, 1, 44_100, 48_000, 144_000, 0, b.FMT_S16)
    assert_tables(res, ref)
    assert rel_err(out[: res.nseg], ref.segments.astype(np.float64)) <= RESAMPLE_TOL


def test_resample_linearity_full_size(ctx):
    """Size-independent property at C2's full window size: the resampler is linear."""
    import torch
    n = 132_300 * 6
    a = synth_pcm(31, n / 44_100, 44_100, 1, np.float32)[:n]
    c = synth_pcm(32, n / 44_100, 44_100, 1, np.float32)[:n]
    plan = b.FrontEndPlan(ctx, 44_100, 1, b.FMT_F32, 48_000, 144_000, 0)
    outs = []
    for x in (a, c, (a + c).astype(np.float32)):
        r = plan.run(x); ctx.sync()
        outs.append(r.torch().cpu().numpy()[: r.nseg].astype(np.float64).copy())
    scale = np.sqrt(np.mean(outs[2] ** 2))
    assert np.abs(outs[0] + outs[1] - outs[2]).max() <= 2e-5 * scale


def test_plan_errors(ctx):
    with pytest.raises(b.BirdaError) as e:
        b.FrontEndPlan(ctx, 48_000, 1, b.FMT_S16, 48_000, 100, 100)
    assert e.value.code == -2
    with pytest.raises(b.BirdaError) as e:
        b.FrontEndPlan(ctx, 44_101, 1, b.FMT_S16, 48_000, 144_000, 0)
    assert e.value.code == -3
    with pytest.raises(b.BirdaError) as e:
        b.FrontEndPlan(ctx, 48_000, 1, 9, 48_000, 144_000, 0)
    assert e.value.code == -4
