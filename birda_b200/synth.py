"""Seeded synthetic audio / logits of SURVEY.md §8(d) (shared by tests and bench.py)."""
from __future__ import annotations

import numpy as np


def synth_pcm(seed: int, seconds: float, rate: int, channels: int = 1, dtype=np.int16, bat: bool = False,
              block_seconds: float = 60.0) -> np.ndarray:
    """Sum of 3-6 chirps/tones at -12 dBFS + Gaussian noise at -30 dBFS, quantised; channel c>0 is
    channel 0 delayed by 7*c frames * 0.8 + independent noise.  Returns interleaved [frames*channels]."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * rate))
    lo, hi = (20_000.0, 110_000.0) if bat else (500.0, 12_000.0)
    hi = min(hi, 0.45 * rate)
    lo = min(lo, hi * 0.25)
    k = int(rng.integers(3, 7))
    f0 = rng.uniform(lo, hi, k); f1 = rng.uniform(lo, hi, k); ph = rng.uniform(0, 2 * np.pi, k)
    amp = 10 ** (-12 / 20) / k
    out = np.empty((n, channels), dtype=dtype)
    blk = max(int(block_seconds * rate), 1)
    for s in range(0, n, blk):
        e = min(n, s + blk)
        t = np.arange(s, e, dtype=np.float64) / rate
        x = np.zeros(e - s)
        for i in range(k):
            # slow linear sweep f0 -> f1 over 10 s, repeating
            tt = np.mod(t, 10.0)
            x += amp * np.sin(ph[i] + 2 * np.pi * (f0[i] * tt + 0.5 * (f1[i] - f0[i]) / 10.0 * tt * tt))
        x += 10 ** (-30 / 20) * rng.standard_normal(e - s)
        for c in range(channels):
            if c == 0:
                y = x
            else:
                y = 0.8 * np.roll(x, 7 * c) + 10 ** (-30 / 20) * rng.standard_normal(e - s)
            if dtype == np.int16:
                out[s:e, c] = np.clip(np.round(y * 32767.0), -32768, 32767).astype(np.int16)
            elif dtype == np.int32:
                out[s:e, c] = np.clip(np.round(y * 2147483647.0), -2147483648, 2147483647).astype(np.int32)
            else:
                out[s:e, c] = y.astype(np.float32)
    return out.reshape(-1)


def pack_s24(samples_i32: np.ndarray) -> np.ndarray:
    """int32 values in [-2^23, 2^23) -> 3-byte little-endian packed PCM (uint8, 3 per sample) as a 24-bit WAV holds it."""
    v = np.asarray(samples_i32, dtype=np.int32).astype(np.uint32)
    out = np.empty((v.size, 3), np.uint8)
    out[:, 0] = v & 0xFF; out[:, 1] = (v >> 8) & 0xFF; out[:, 2] = (v >> 16) & 0xFF
    return out.reshape(-1)


def synth_pcm24(seed: int, seconds: float, rate: int, channels: int = 1) -> np.ndarray:
    """The synthetic signal quantised to 24 bits, packed (interleaved, 3 bytes per sample)."""
    x = synth_pcm(seed, seconds, rate, channels, np.int32)
    return pack_s24(x >> 8)


def synth_logits(seed: int, rows: int, classes: int, adversarial: bool = True) -> np.ndarray:
    """N(-6, 2^2) background with 0-8 planted N(2,1) values per row, plus adversarial rows
    (exact ties, values exactly at logit(0.1), all below threshold, >= 6 above)."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, classes)) * 2.0 - 6.0).astype(np.float32)
    for r in range(rows):
        m = int(rng.integers(0, 9))
        idx = rng.choice(classes, m, replace=False)
        x[r, idx] = rng.standard_normal(m).astype(np.float32) + 2.0
    if adversarial and rows >= 6:
        x[0, :] = -20.0                                   # nothing above threshold
        x[1, :] = -20.0; x[1, [5 % classes, 17 % classes, 300 % classes, 301 % classes, 4000 % classes, classes - 1, 0]] = 1.25   # exact ties
        x[2, :] = -20.0; x[2, 10] = np.float32(np.log(0.1 / 0.9))                        # at the threshold
        x[3, :] = -20.0; x[3, rng.choice(classes, 9, replace=False)] = rng.uniform(0, 3, 9).astype(np.float32)
        x[4, :] = 30.0                                    # everything saturates to 1.0: lowest indices win
        x[5, :] = -20.0; x[5, classes - 3:] = 5.0         # ties at the end of the row
    return x


def write_wav(path: str, pcm: np.ndarray, rate: int, channels: int) -> None:
    """Plain RIFF/WAVE file of interleaved int16 / int32 / float32 samples (benches and tools)."""
    import struct
    pcm = np.ascontiguousarray(pcm)
    tag, bits = {np.dtype(np.int16): (1, 16), np.dtype(np.int32): (1, 32), np.dtype(np.float32): (3, 32)}[pcm.dtype]
    raw = pcm.tobytes()
    fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(raw)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"data" + struct.pack("<I", len(raw)))
        f.write(raw)
