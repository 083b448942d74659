"""K5 — spectrogram prefix (STFT power + tcgen05 mel projection) through the C ABI against the float64 oracle.
The reference pins none of this layer's parameters: parity here is against our own statement."""
import numpy as np
import pytest

import birda_b200 as b
from birda_b200.synth import synth_pcm
from oracle import melspec as om

pytestmark = pytest.mark.gpu
REL_TOL = 2e-5     # of the row's largest mel energy (f32 FFT + three-product tf32 split; measured ~2e-6)


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0)
    yield c
    c.close()


def run_gpu(ctx, seg, n_fft, hop, n_frames, window, mw, **kw):
    import torch
    d = torch.from_numpy(seg).cuda()
    out = torch.full((seg.shape[0], mw.shape[0], n_frames), float("nan"), device="cuda")
    torch.cuda.synchronize()          # the fill runs on torch's stream, the kernels on the context's own stream
    ms = b.MelSpec(ctx, n_fft, hop, n_frames, window, mw, **kw)
    ms.run(d.data_ptr(), seg.shape[0], seg.shape[1], out.data_ptr())
    ctx.sync()
    info = ms.info()
    ms.close()
    return out.cpu().numpy(), info


def windows(seed, rows, samples, rate):
    pcm = synth_pcm(seed, rows * samples / rate, rate, 1).astype(np.float32) / 32768.0
    return np.ascontiguousarray(pcm[: rows * samples].reshape(rows, samples))


@pytest.mark.parametrize("n_fft,hop,n_frames,n_mels,fmin,fmax", [
    (2048, 278, 511, 96, 0.0, 3000.0),          # BirdNET-v2.4-like low band (narrow support: 129 bins)
    (1024, 280, 511, 96, 500.0, 15000.0),       # BirdNET-v2.4-like high band
    (512, 160, 300, 64, 60.0, 16000.0),         # wide support, other sizes
    (2048, 512, 281, 128, 0.0, 24000.0),        # full band: every bin, K = 1025 -> 1056
])
def test_melspec_linear(ctx, n_fft, hop, n_frames, n_mels, fmin, fmax):
    rate, samples, rows = 48_000, 144_000, 5
    seg = windows(11, rows, samples, rate)
    w, mw = om.hann(n_fft), om.mel_filterbank(n_mels, n_fft, rate, fmin, fmax)
    got, (lo, nb, kpad) = run_gpu(ctx, seg, n_fft, hop, n_frames, w, mw)
    ref = om.melspec(seg, n_fft, hop, n_frames, w, mw)
    assert kpad % 32 == 0 and kpad >= nb and np.all(mw[:, :lo] == 0) and np.all(mw[:, lo + nb:] == 0)
    assert np.isfinite(got).all()
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    err = (np.abs(got - ref) / scale).max()
    assert err <= REL_TOL, err


def test_melspec_log_modes_and_magnitude(ctx):
    rate, samples, rows, n_fft, hop, n_frames = 32_000, 160_000, 3, 1024, 320, 500
    seg = windows(12, rows, samples, rate)
    w, mw = om.hann(n_fft), om.mel_filterbank(64, n_fft, rate, 100.0, 14000.0)
    for kw in (dict(power=1.0), dict(log_mode=1, log_eps=1e-6), dict(log_mode=2, log_eps=1e-10), dict(power=1.5)):
        got, _ = run_gpu(ctx, seg, n_fft, hop, n_frames, w, mw, **kw)
        ref = om.melspec(seg, n_fft, hop, n_frames, w, mw, **kw)
        tol = 2e-5 * np.abs(ref).max() if kw.get("log_mode", 0) == 0 else 2e-3      # log of tiny energies amplifies f32 noise
        assert np.abs(got - ref).max() <= tol, (kw, np.abs(got - ref).max())


def test_melspec_many_rows_chunked(ctx):
    """More rows than one chunk of the power buffer holds (two STFT / GEMM launch pairs), ragged tail tile."""
    rate, samples, rows, n_fft, hop, n_frames = 48_000, 24_000, 310, 1024, 47, 511
    seg = windows(13, rows, samples, rate)
    w, mw = om.hann(n_fft), om.mel_filterbank(80, n_fft, rate, 0.0, 24000.0)
    l0 = ctx.kernel_launches
    got, _ = run_gpu(ctx, seg, n_fft, hop, n_frames, w, mw)
    assert ctx.kernel_launches - l0 >= 4
    ref = om.melspec(seg, n_fft, hop, n_frames, w, mw)
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got - ref) / scale).max() <= REL_TOL


@pytest.mark.parametrize("n_fft,hop,samples,n_frames,rows", [
    (2048, 277, 47_999, 190, 3),     # odd hop and odd row length: unaligned frame starts, ragged tail, frames wholly past the end
    (1024, 1, 3_000, 64, 1),         # hop 1, a single short row
    (256, 100, 5_000, 70, 2),        # single-stream runtime-plan STFT kernel (n_fft without a two-stream plan), frames past the end
    (4096, 512, 20_000, 9, 5),       # largest frame length
])
def test_melspec_edge_shapes(ctx, n_fft, hop, samples, n_frames, rows):
    rate = 48_000
    rng = np.random.default_rng(n_fft + hop)
    seg = (rng.standard_normal((rows, samples)) * 0.2).astype(np.float32)
    w, mw = om.hann(n_fft), om.mel_filterbank(32, n_fft, rate, 300.0, 20_000.0)
    got, _ = run_gpu(ctx, seg, n_fft, hop, n_frames, w, mw)
    ref = om.melspec(seg, n_fft, hop, n_frames, w, mw)
    assert np.isfinite(got).all()
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got - ref) / scale).max() <= REL_TOL
    past = [t for t in range(n_frames) if t * hop >= samples]
    if past:
        assert not got[:, :, past].any()          # frames that start past the end of the row are silence


def test_melspec_errors(ctx):
    w, mw = om.hann(1024), om.mel_filterbank(64, 1024, 48_000, 0.0, 24000.0)
    with pytest.raises(b.BirdaError):
        b.MelSpec(ctx, 1000, 100, 10, np.ones(1000, np.float32), np.ones((64, 501), np.float32))   # not a power of two
    with pytest.raises(b.BirdaError):
        b.MelSpec(ctx, 1024, 100, 10, w, mw[:60])                                                  # n_mels % 16
    with pytest.raises(b.BirdaError):
        b.MelSpec(ctx, 1024, 0, 10, w, mw)
