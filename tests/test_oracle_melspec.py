"""The float64 statement of the spectrogram prefix (oracle/melspec.py) against a literal DFT and known signals."""
import numpy as np

from oracle import melspec as om


def test_single_tone_lands_in_its_filter():
    rate, n_fft, hop, n_frames = 48_000, 1024, 256, 20
    t = np.arange(n_fft + hop * (n_frames - 1)) / rate
    x = np.sin(2 * np.pi * 3000.0 * t)[None, :].astype(np.float32)
    mw = om.mel_filterbank(64, n_fft, rate, 0.0, 24000.0)
    m = om.melspec(x, n_fft, hop, n_frames, om.hann(n_fft), mw)
    assert m.shape == (1, 64, n_frames)
    k = int(round(3000.0 / rate * n_fft))
    best = int(np.argmax(mw[:, k]))
    assert np.all(np.abs(np.argmax(m[0], axis=0) - best) <= 1)
    assert np.allclose(m[0, :, 0], m[0, :, -1], rtol=0.2)          # stationary signal


def test_against_literal_dft_and_zero_padding():
    rng = np.random.default_rng(3)
    n_fft, hop, n_frames = 256, 100, 7
    x = rng.standard_normal((2, 700)).astype(np.float32)            # last frames run past the end: zero padded
    w = om.hann(n_fft)
    mw = om.mel_filterbank(16, n_fft, 16_000, 100.0, 7000.0)
    got = om.melspec(x, n_fft, hop, n_frames, w, mw, power=1.0)
    k = np.arange(n_fft // 2 + 1)[:, None] * np.arange(n_fft)[None, :]
    F = np.exp(-2j * np.pi * k / n_fft)
    for r in range(2):
        for t in range(n_frames):
            fr = np.zeros(n_fft)
            seg = x[r, t * hop: t * hop + n_fft]
            fr[: len(seg)] = seg
            mag = np.abs(F @ (fr * w.astype(np.float64)))
            assert np.allclose(got[r, :, t], mw.astype(np.float64) @ mag, rtol=1e-9, atol=1e-12)


def test_filterbank_shape_and_support():
    mw = om.mel_filterbank(96, 2048, 48_000, 0.0, 3000.0)
    assert mw.shape == (96, 1025) and mw.dtype == np.float32 and (mw >= 0).all()
    nz = np.nonzero(mw.any(axis=0))[0]
    assert nz[0] >= 1 and nz[-1] <= int(3000.0 / 48_000 * 2048) + 1
