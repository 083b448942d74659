"""Oracle (TEST INFRASTRUCTURE, not product code) for the spectrogram prefix K5.

The reference has no host code for this step: the STFT / mel layers live inside the ONNX graphs
(manifests/Perch-v2-Models.models.json:15,46; BirdNET v2.4's in-graph spectrogram) and nothing under
/root/reference pins frame length, hop, mel edges or scaling -> "parity unpinned" against the reference.
This file defines the layer the kernels implement, in float64, so the CUDA path has an exact statement to
be checked against:  frame t = samples [t*hop, t*hop + n_fft) (zero past the end) * window -> rfft ->
|X|^power -> mel_weights -> optional log scaling -> [rows, n_mels, n_frames].
Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this package."""
from __future__ import annotations

import numpy as np


def hann(n_fft: int) -> np.ndarray:
    """Periodic Hann window (tf.signal.hann_window default), float32."""
    n = np.arange(n_fft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(np.float32)


def mel_filterbank(n_mels: int, n_fft: int, sample_rate: int, fmin: float, fmax: float) -> np.ndarray:
    """Triangular HTK-mel filters [n_mels, n_fft/2 + 1] (tf.signal.linear_to_mel_weight_matrix layout, transposed)."""
    def hz_to_mel(f):
        return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    bins = n_fft // 2 + 1
    freqs = np.linspace(0.0, sample_rate / 2.0, bins)
    edges = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
    m = hz_to_mel(freqs)
    w = np.zeros((n_mels, bins), dtype=np.float64)
    for i in range(n_mels):
        lo, ce, hi = edges[i], edges[i + 1], edges[i + 2]
        up = (m - lo) / (ce - lo)
        dn = (hi - m) / (hi - ce)
        w[i] = np.maximum(0.0, np.minimum(up, dn))
    w[:, 0] = 0.0
    return w.astype(np.float32)


def melspec(segments: np.ndarray, n_fft: int, hop: int, n_frames: int, window: np.ndarray, mel_weights: np.ndarray,
            power: float = 2.0, log_mode: int = 0, log_eps: float = 1e-6) -> np.ndarray:
    """float64 statement of the layer; segments [rows, samples] -> [rows, n_mels, n_frames]."""
    x = np.asarray(segments, dtype=np.float64)
    rows, samples = x.shape
    need = (n_frames - 1) * hop + n_fft
    if need > samples:
        x = np.concatenate([x, np.zeros((rows, need - samples))], axis=1)
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    w64 = np.asarray(window, dtype=np.float32).astype(np.float64)
    mw = np.asarray(mel_weights, dtype=np.float32).astype(np.float64)
    out = np.empty((rows, mw.shape[0], n_frames), dtype=np.float64)
    for r in range(rows):
        fr = x[r][idx] * w64[None, :]
        mag2 = np.abs(np.fft.rfft(fr, axis=1)) ** 2
        p = mag2 if power == 2.0 else mag2 ** (0.5 * power)
        m = p @ mw.T
        if log_mode == 1:
            m = np.log(m + log_eps)
        elif log_mode == 2:
            m = 10.0 * np.log10(np.maximum(m, log_eps))
        out[r] = m.T
    return out
