"""Front-end oracle: PCM -> mono f32 -> windows -> per-window FFT resample -> packed segments.

Test infrastructure only (see oracle/__init__.py).  numpy restatement, f32-faithful
where the reference computes in f32; an optional f64 mode of the FFT core exists to
measure the f32 round-off floor.  Paths cited are relative to /root/reference.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

try:  # scipy's pocketfft computes float32 transforms natively in single precision
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = np.fft

from . import rules

f32 = np.float32

FMT_S16, FMT_S32, FMT_F32, FMT_S24 = 1, 2, 3, 4   # mirrors bb_sample_fmt in include/birda_b200.h


def s24_to_s32(packed: np.ndarray) -> np.ndarray:
    """3-byte little-endian PCM -> the S32 buffer symphonia's PCM decoder hands to append_samples for a 24-bit
    stream: ``sample << 8`` (symphonia-codec-pcm, not vendored under /root/reference: stated from the crate's
    published behaviour; the reference then takes the S32 arm, src/audio/decode.rs:386-402)."""
    b = np.asarray(packed, dtype=np.uint8).reshape(-1, 3).astype(np.uint32)
    return ((b[:, 0] << np.uint32(8)) | (b[:, 1] << np.uint32(16)) | (b[:, 2] << np.uint32(24))).astype(np.uint32).view(np.int32)


# --------------------------------------------------------------------------- A1
def to_mono_f32(pcm: np.ndarray, channels: int) -> np.ndarray:
    """Sample conversion + downmix.  src/audio/decode.rs:353-411.

    ``pcm``: interleaved frames, shape [frames*channels] or [frames, channels], dtype
    int16 / int32 / float32.  Mono: straight convert.  Multi-channel: ``sum = 0f32``,
    left-to-right f32 adds of the converted samples, one f32 divide by ``channels``."""
    a = np.asarray(pcm)
    if a.dtype == np.uint8:                                           # packed 24-bit PCM (FMT_S24)
        a = s24_to_s32(a)
    if a.dtype == np.int16:
        conv = lambda x: x.astype(np.float32) / f32(32768.0)          # f32::from(s) / 32768.0
    elif a.dtype == np.int32:
        conv = lambda x: x.astype(np.float32) / f32(2147483648.0)     # s as f32 / 2^31 (RNE cast)
    elif a.dtype == np.float32:
        conv = lambda x: x
    else:
        raise TypeError("unsupported sample format (the reference silently drops it, decode.rs:407-409)")
    a = a.reshape(-1, channels)
    if channels == 1:
        return np.ascontiguousarray(conv(a[:, 0]), dtype=np.float32)
    s = np.zeros(a.shape[0], dtype=np.float32)
    for ch in range(channels):
        s = (s + conv(a[:, ch])).astype(np.float32)
    return (s / f32(channels)).astype(np.float32)


# --------------------------------------------------------------------------- A4
@dataclass
class ResamplerPlan:
    """Block sizes and filter spectrum of ``rubato::Fft::<f32>::new(from, to, 1024, 1, FixedSync::Both)``.

    Third-party (rubato 4.0.0; not vendored under /root/reference) — restated from the
    published algorithm (synchro.rs / sinc.rs / windows.rs of rubato <= 1.0); call site
    src/audio/resample.rs:19-25; block-size evidence src/audio/resample.rs:313-316.
    SAMPLE-LEVEL PARITY UNPINNED."""
    from_rate: int
    to_rate: int
    n_in: int
    n_out: int
    n_keep: int            # spectrum bins carried from the forward to the inverse transform
    cutoff: np.float32
    taps: np.ndarray       # [n_in] f32 time-domain filter, already / (2*n_in)
    filt_f: np.ndarray     # [n_in+1] complex64 = rfft(taps zero-padded to 2*n_in)


def blackman_harris2_periodic(n: int) -> np.ndarray:
    """Squared 4-term Blackman-Harris, periodic form (denominator n), in f32."""
    x = np.arange(n, dtype=np.float32)
    npf = f32(n)
    pi = f32(np.pi)
    a, b, c, d = f32(0.35875), f32(0.48829), f32(0.14128), f32(0.01168)
    w = (a - b * np.cos(f32(2.0) * pi * x / npf)
         + c * np.cos(f32(4.0) * pi * x / npf)
         - d * np.cos(f32(6.0) * pi * x / npf)).astype(np.float32)
    return (w * w).astype(np.float32)


def make_plan(from_rate: int, to_rate: int, chunk_size: int = 1024) -> ResamplerPlan:
    g = math.gcd(from_rate, to_rate)
    min_in = from_rate // g
    k = -(-chunk_size // min_in)                 # ceil(1024 / (from/g))
    n_in = k * from_rate // g
    n_out = k * to_rate // g
    cutoff = f32(0.4) ** f32(f32(16.0) / f32(n_in))
    if n_in > n_out:
        cutoff = f32(f32(cutoff * f32(n_out)) / f32(n_in))
    cutoff = f32(cutoff)
    w = blackman_harris2_periodic(n_in)
    x = (np.arange(n_in, dtype=np.float32) - f32(n_in // 2)) * cutoff
    xpi = (x * f32(np.pi)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(x == 0, f32(1.0), np.sin(xpi) / xpi).astype(np.float32)
    y = (w * s).astype(np.float32)
    total = f32(0.0)
    for v in y:                                   # sequential f32 sum, as the crate's loop
        total = f32(total + v)
    taps = ((y / total).astype(np.float32) / f32(2 * n_in)).astype(np.float32)
    padded = np.zeros(2 * n_in, dtype=np.float32)
    padded[:n_in] = taps
    filt_f = _fft.rfft(padded).astype(np.complex64)
    n_keep = n_in + 1 if n_in < n_out else n_out
    return ResamplerPlan(from_rate, to_rate, n_in, n_out, n_keep, cutoff, taps, filt_f)


def _resample_unit(plan: ResamplerPlan, x: np.ndarray, carry: np.ndarray, dtype) -> np.ndarray:
    """One ``process()`` call: 2N_in real FFT, filter, re-bin, 2N_out inverse, overlap-add."""
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    buf = np.zeros(2 * plan.n_in, dtype=dtype)
    buf[: plan.n_in] = x
    xf = _fft.rfft(buf).astype(cdt)
    yf = np.zeros(plan.n_out + 1, dtype=cdt)
    yf[: plan.n_keep] = xf[: plan.n_keep] * plan.filt_f[: plan.n_keep].astype(cdt)
    # realfft's inverse ignores the imaginary part of the DC and Nyquist bins and is
    # unnormalised; scipy's irfft divides by the length, so multiply it back.
    y = _fft.irfft(yf, n=2 * plan.n_out).astype(dtype) * dtype(2 * plan.n_out)
    out = (y[: plan.n_out] + carry).astype(dtype)
    carry[:] = y[plan.n_out:]
    return out


def resample(samples: np.ndarray, from_rate: int, to_rate: int,
             plan: Optional[ResamplerPlan] = None, precision: str = "f32") -> np.ndarray:
    """``resample`` / ``resample_chunk``.  src/audio/resample.rs:10-105.

    A brand-new resampler per call (the carry starts at zero and the last carry is
    discarded); full blocks, then one zero-padded block of which only
    ``ceil(remaining*to/from)`` output frames are taken."""
    samples = np.asarray(samples, dtype=np.float32)
    if from_rate == to_rate:
        return samples
    if plan is None:
        plan = make_plan(from_rate, to_rate)
    dtype = np.float32 if precision == "f32" else np.float64
    n_in, n_out = plan.n_in, plan.n_out
    carry = np.zeros(n_out, dtype=dtype)
    outs: List[np.ndarray] = []
    pos = 0
    while pos + n_in <= samples.size:
        outs.append(_resample_unit(plan, samples[pos: pos + n_in].astype(dtype), carry, dtype))
        pos += n_in
    if pos < samples.size:
        remaining = samples.size - pos
        padded = np.zeros(n_in, dtype=dtype)
        padded[:remaining] = samples[pos:]
        block = _resample_unit(plan, padded, carry, dtype)
        output_frames = math.ceil(float(remaining) * float(to_rate) / float(from_rate))
        outs.append(block[: min(output_frames, block.size)])
    if not outs:
        return np.zeros(0, dtype=np.float32)
    return np.concatenate(outs).astype(np.float32)


def resample_batched(samples: np.ndarray, from_rate: int, to_rate: int, plan: Optional[ResamplerPlan] = None) -> np.ndarray:
    """``resample`` with all blocks of the call transformed at once (one rfft / irfft over a [blocks, 2N] array, f32) —
    the same arithmetic as the block loop up to the order of the overlap-add sums.  This is the form the CPU BASELINE
    of bench.py times: it removes the Python per-block overhead, which the reference (compiled Rust) does not pay, so
    the baseline is not flattered.  Held to ``resample`` at 2e-6 by tests/test_oracle_resample.py."""
    samples = np.asarray(samples, dtype=np.float32)
    if from_rate == to_rate:
        return samples
    if plan is None:
        plan = make_plan(from_rate, to_rate)
    n_in, n_out = plan.n_in, plan.n_out
    nblk = -(-samples.size // n_in)
    if nblk == 0:
        return np.zeros(0, dtype=np.float32)
    buf = np.zeros((nblk, 2 * n_in), dtype=np.float32)
    flat = np.zeros(nblk * n_in, dtype=np.float32)
    flat[: samples.size] = samples
    buf[:, :n_in] = flat.reshape(nblk, n_in)
    xf = _fft.rfft(buf, axis=1)
    yf = np.zeros((nblk, n_out + 1), dtype=np.complex64)
    yf[:, : plan.n_keep] = xf[:, : plan.n_keep] * plan.filt_f[: plan.n_keep]
    yf[:, 0] = yf[:, 0].real                      # realfft ignores the imaginary part of DC and Nyquist
    yf[:, -1] = yf[:, -1].real
    y = (_fft.irfft(yf, n=2 * n_out, axis=1) * np.float32(2 * n_out)).astype(np.float32)
    out = y[:, :n_out].copy()
    out[1:] += y[:-1, n_out:]                     # overlap-add: the second half of block b lands on block b + 1
    return out.reshape(-1)[: resampled_len(samples.size, plan)]


def resampled_len(src_len: int, plan: ResamplerPlan) -> int:
    """Length ``resample`` returns for ``src_len`` input samples (before the resize)."""
    nfull, rem = divmod(src_len, plan.n_in)
    n = nfull * plan.n_out
    if rem:
        n += min(math.ceil(float(rem) * float(plan.to_rate) / float(plan.from_rate)), plan.n_out)
    return n


# --------------------------------------------------------------------------- A2+A3+A4
@dataclass
class FrontEndResult:
    segments: np.ndarray        # [nseg, segment_samples] f32  (AudioChunk.samples)
    start_sample: np.ndarray    # [nseg] uint64               (RawSegment.start_sample)
    start_time: np.ndarray      # [nseg] f32
    end_time: np.ndarray        # [nseg] f32
    src_seg: int
    src_ovl: int


def decode_and_stream(pcm: np.ndarray, channels: int, source_rate: int, target_rate: int,
                      segment_samples: int, overlap_samples: int,
                      precision: str = "f32", only: Optional[range] = None, batched: bool = False) -> FrontEndResult:
    """Everything the decode thread does between decoded PCM and ``tx.send``.

    src/pipeline/processor.rs:49-108 over src/audio/decode.rs:150-202 and
    src/audio/resample.rs:97-105.  ``only`` restricts which segment rows are
    *computed* (the tables are always complete) so large workloads stay cheap."""
    mono = to_mono_f32(pcm, channels)
    src_seg, src_ovl = rules.source_window(segment_samples, overlap_samples, source_rate, target_rate)
    table = rules.next_segment_table(mono.size, src_seg, src_ovl)
    nseg = len(table)
    plan = make_plan(source_rate, target_rate) if source_rate != target_rate else None
    segs = np.zeros((nseg, segment_samples), dtype=np.float32)
    st = np.zeros(nseg, dtype=np.float32)
    et = np.zeros(nseg, dtype=np.float32)
    ss = np.zeros(nseg, dtype=np.uint64)
    rows = range(nseg) if only is None else only
    for i, w in enumerate(table):
        ss[i] = w.start_sample
        st[i], et[i] = rules.chunk_times(w.start_sample, source_rate, segment_samples, target_rate)
    for i in rows:
        w = table[i]
        raw = np.zeros(src_seg, dtype=np.float32)
        raw[: w.take] = mono[w.start_sample: w.start_sample + w.take]
        out = resample_batched(raw, source_rate, target_rate, plan) if batched else resample(raw, source_rate, target_rate, plan, precision)
        n = min(out.size, segment_samples)            # samples.resize(segment_samples, 0.0)
        segs[i, :n] = out[:n]
    return FrontEndResult(segs, ss, st, et, src_seg, src_ovl)
