"""Python mirror of the C ABI (tests, bench and Python hosts).  No compute happens here."""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import BirdaError, MelSpecCfg, PostCfg, check, lib

FMT_S16, FMT_S32, FMT_F32, FMT_S24 = 1, 2, 3, 4      # FMT_S24: 3-byte packed little-endian (uint8 arrays)
ACT_NONE, ACT_SIGMOID, ACT_SOFTMAX = 0, 1, 2
_NP_FMT = {np.dtype(np.int16): FMT_S16, np.dtype(np.int32): FMT_S32, np.dtype(np.float32): FMT_F32, np.dtype(np.uint8): FMT_S24}


class rules:
    """Host rules (pure arithmetic in the library; no GPU).  Names follow the reference."""

    @staticmethod
    def segment_samples(segment_duration: float, overlap: float, target_rate: int, bat_mode: bool = False) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.bb_rule_segment_samples(segment_duration, overlap, target_rate, int(bat_mode), C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def source_window(seg: int, ovl: int, src_rate: int, tgt_rate: int) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.bb_rule_source_window(seg, ovl, src_rate, tgt_rate, C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def segment_count(total_frames: int, src_seg: int, src_ovl: int) -> int:
        n = C.c_uint64()
        check(lib.bb_rule_segment_count(total_frames, src_seg, src_ovl, C.byref(n)))
        return n.value

    @staticmethod
    def segment_table(total_frames: int, src_seg: int, src_ovl: int, first: int = 0, capacity: Optional[int] = None):
        if capacity is None:
            capacity = rules.segment_count(total_frames, src_seg, src_ovl)
        st = np.zeros(max(capacity, 1), np.uint64)
        tk = np.zeros(max(capacity, 1), np.uint64)
        w = C.c_uint64()
        check(lib.bb_rule_segment_table(total_frames, src_seg, src_ovl, first, capacity,
                                        st.ctypes.data_as(_lib.u64p), tk.ctypes.data_as(_lib.u64p), C.byref(w)))
        return st[: w.value], tk[: w.value]

    @staticmethod
    def chunk_times(start_sample: int, src_rate: int, seg: int, tgt_rate: int) -> Tuple[np.float32, np.float32]:
        a, b = C.c_float(), C.c_float()
        check(lib.bb_rule_chunk_times(start_sample, src_rate, seg, tgt_rate, C.byref(a), C.byref(b)))
        return np.float32(a.value), np.float32(b.value)

    @staticmethod
    def estimate_segment_count(duration: Optional[float], segment_duration: float, overlap: float) -> Optional[int]:
        e = C.c_int64()
        check(lib.bb_rule_estimate_segment_count(0.0 if duration is None else duration, int(duration is not None),
                                                 segment_duration, overlap, C.byref(e)))
        return None if e.value < 0 else e.value

    @staticmethod
    def effective_batch_size(batch: int, estimate: Optional[int]) -> int:
        return lib.bb_rule_effective_batch_size(batch, -1 if estimate is None else estimate)

    @staticmethod
    def resampler_blocks(src_rate: int, tgt_rate: int):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_float()
        check(lib.bb_rule_resampler_blocks(src_rate, tgt_rate, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, np.float32(d.value)

    @staticmethod
    def resampler_taps(src_rate: int, tgt_rate: int) -> np.ndarray:
        n_in = rules.resampler_blocks(src_rate, tgt_rate)[0]
        t = np.zeros(n_in, np.float32)
        check(lib.bb_rule_resampler_taps(src_rate, tgt_rate, t.ctypes.data_as(_lib.f32p), n_in))
        return t

    @staticmethod
    def resampled_len(src_len: int, src_rate: int, tgt_rate: int) -> int:
        n = C.c_uint64()
        check(lib.bb_rule_resampled_len(src_len, src_rate, tgt_rate, C.byref(n)))
        return n.value

    @staticmethod
    def inference_timeout_secs(env_value: Optional[str]) -> int:
        """BIRDA_INFERENCE_TIMEOUT parsing (src/pipeline/processor.rs:194-211)."""
        return lib.bb_rule_inference_timeout_secs(None if env_value is None else env_value.encode())

    @staticmethod
    def scientific_name(label: str) -> str:
        """src/inference/geomodel.rs:28-33."""
        raw = label.encode()
        return raw[: lib.bb_rule_scientific_name_len(raw)].decode()

    date_to_week = staticmethod(lambda m, d: lib.bb_rule_date_to_week(m, d))
    week_to_start_day = staticmethod(lambda w: lib.bb_rule_week_to_start_day(w))

    @staticmethod
    def day_of_year_to_date(doy: int) -> Tuple[int, int]:
        a, b = C.c_uint32(), C.c_uint32()
        lib.bb_rule_day_of_year_to_date(doy, C.byref(a), C.byref(b))
        return a.value, b.value


def mask_build(classifier_labels, geomodel_labels, scores) -> Tuple[np.ndarray, int, int]:
    """Dense range mask in the classifier's label space (``bb_mask_build``): ``scores`` is a sequence of
    (geomodel species label, score).  Returns (mask [C] f32 with NaN = no geomodel entry, mapped, unmatched)."""
    def arr(strings):
        enc = [s.encode() for s in strings]
        return (C.c_char_p * max(len(enc), 1))(*enc), enc
    cl, _k1 = arr(classifier_labels)
    gl, _k2 = arr(geomodel_labels)
    sp, _k3 = arr([s for s, _ in scores])
    sv = np.asarray([v for _, v in scores], dtype=np.float32)
    mask = np.zeros(max(len(classifier_labels), 1), np.float32)
    mapped, unmatched = C.c_uint32(), C.c_uint32()
    check(lib.bb_mask_build(cl, len(classifier_labels), gl, len(geomodel_labels), sp, sv.ctypes.data_as(_lib.f32p), len(sv),
                            mask.ctypes.data_as(_lib.f32p), C.byref(mapped), C.byref(unmatched)))
    return mask[: len(classifier_labels)], mapped.value, unmatched.value


def flac_probe(path_or_bytes):
    """STREAMINFO of a FLAC file (path) or stream (bytes): ``_lib.FlacInfo``."""
    info = _lib.FlacInfo()
    if isinstance(path_or_bytes, (bytes, bytearray)):
        raw = bytes(path_or_bytes)
        check(lib.bb_flac_probe_bytes(raw, len(raw), C.byref(info)))
    else:
        check(lib.bb_flac_probe(str(path_or_bytes).encode(), C.byref(info)))
    return info


class FlacDecoder:
    """``bb_flac``: the compressed file goes to the GPU, one thread per frame decodes it into the interleaved PCM the
    front end consumes.  ``decode(bytes)`` -> (device pointer, frames, info); the buffer is valid until the next decode."""

    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        self._h = C.c_void_p()
        check(lib.bb_flac_create(ctx.handle, C.byref(self._h)), ctx.handle)

    def decode(self, data: bytes):
        info = flac_probe(data)
        ptr, frames = C.c_void_p(), C.c_uint64()
        check(lib.bb_flac_decode(self._h, data, len(data), C.byref(info), C.byref(ptr), C.byref(frames)), self.ctx.handle)
        return ptr.value or 0, int(frames.value), info

    def decode_to_numpy(self, data: bytes) -> Tuple[np.ndarray, "_lib.FlacInfo"]:
        """Decoded PCM copied back to the host: int16 / packed-24 (uint8) / int32, interleaved (tests)."""
        ptr, frames, info = self.decode(data)
        dt, per = {FMT_S16: (np.int16, 1), FMT_S24: (np.uint8, 3), FMT_S32: (np.int32, 1)}[info.fmt]
        out = np.zeros(frames * info.channels * per, dt)
        if out.size:
            check(lib.bb_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes), self.ctx.handle)
            self.ctx.sync()
        return out, info

    def close(self):
        if self._h:
            lib.bb_flac_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StandIn:
    """``bb_standin``: the library's stand-in classifier (NOT a model) — [B, samples] windows -> [B, classes] logits on
    the device, as a native ``bb_classify_fn`` for NativePipeline / NativePool and the benches.  ``stream``: queue on
    that CUDA stream asynchronously (a pipeline on its own context); default: own stream, finished on return."""

    def __init__(self, device: int, samples: int, classes: int, max_batch: int, seed: int = 0, stream=None):
        self.device, self.samples, self.classes, self.max_batch = device, samples, classes, max_batch
        self.handle = C.c_void_p()
        check(lib.bb_standin_create(device, samples, classes, max_batch, seed, C.byref(self.handle)))
        if stream is not None:
            lib.bb_standin_use_stream(self.handle, C.c_void_p(stream), 1)
        self.fn = C.cast(lib.bb_standin_classify, _lib.CLASSIFY_FN)

    def weights(self) -> Tuple[np.ndarray, np.ndarray]:
        W = np.zeros((48, self.classes), np.float32); b = np.zeros(self.classes, np.float32)
        check(lib.bb_standin_weights(self.handle, W.ctypes.data_as(_lib.f32p), b.ctypes.data_as(_lib.f32p)))
        return W, b

    @property
    def launches(self) -> int:
        return int(lib.bb_standin_launches(self.handle))

    def __call__(self, x):
        """torch tensor [B, samples] on the device -> torch tensor [B, classes] (a copy; tests)."""
        import torch

        from .pipeline import _DevView
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.samples
        torch.cuda.current_stream(x.device).synchronize()
        ds, nc = C.c_void_p(), C.c_uint32()
        rc = lib.bb_standin_classify(self.handle, C.c_void_p(x.data_ptr()), x.shape[0], x.shape[1], C.byref(ds), C.byref(nc))
        if rc != 0:
            raise BirdaError(-7, f"stand-in classifier failed ({rc})")
        torch.cuda.synchronize(x.device)
        return torch.as_tensor(_DevView(ds.value, (x.shape[0], nc.value)), device=x.device).clone()

    def close(self):
        if self.handle:
            lib.bb_standin_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Watchdog:
    """``bb_watchdog``: fires ``on_fire(timeout_secs, batch_size)`` unless cancelled within ``timeout_ms``
    (src/gpu/watchdog.rs:22-66).  ``on_fire=None`` is the reference's behaviour: message + exit(1)."""

    def __init__(self, timeout_ms: int, batch_size: int, on_fire=None):
        self._cb = _lib.WATCHDOG_FN(lambda user, secs, batch: on_fire(secs, batch)) if on_fire else C.cast(None, _lib.WATCHDOG_FN)
        self._h = C.c_void_p()
        check(lib.bb_watchdog_start(timeout_ms, batch_size, self._cb, None, C.byref(self._h)))

    def cancel(self):
        if self._h:
            lib.bb_watchdog_cancel(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.cancel()

    def __del__(self):
        try:
            self.cancel()
        except Exception:
            pass


def wav_probe(path: str) -> _lib.WavInfo:
    """Header of a RIFF/RF64 WAVE file (rate, channels, sample format, frames)."""
    info = _lib.WavInfo()
    check(lib.bb_wav_probe(path.encode(), C.byref(info)))
    return info


def wav_read(path: str, info: _lib.WavInfo, first_frame: int = 0, frames: Optional[int] = None) -> np.ndarray:
    """Interleaved frames [first_frame, first_frame+frames) as int16 / int32 / float32 (no conversion); 24-bit
    files come back as the packed bytes (uint8, 3 per sample)."""
    n = info.frames - first_frame if frames is None else frames
    dt = {FMT_S16: np.int16, FMT_S32: np.int32, FMT_F32: np.float32, FMT_S24: np.uint8}[info.fmt]
    out = np.empty(n * info.channels * (3 if info.fmt == FMT_S24 else 1), dtype=dt)
    check(lib.bb_wav_read(path.encode(), C.byref(info), first_frame, n, out.ctypes.data_as(C.c_void_p)))
    return out


def device_count() -> int:
    n = C.c_int32()
    rc = lib.bb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class Context:
    """One GPU + one stream (``bb_ctx``).  ``stream``: an existing cudaStream_t handle (int),
    e.g. ``torch.cuda.current_stream().cuda_stream``, so torch events time the kernels."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        if stream is None:
            check(lib.bb_ctx_create(device, C.byref(self._h)))
        else:
            check(lib.bb_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(self._h)))
        self.device = device
        self._plans = weakref.WeakSet()       # plans must be destroyed before their context

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        """The context's cudaStream_t as an integer (``bb_ctx_stream``)."""
        return int(lib.bb_ctx_stream(self._h) or 0)

    def sync(self):
        check(lib.bb_sync(self._h), self._h)

    @property
    def kernel_launches(self) -> int:
        return lib.bb_ctx_kernel_launches(self._h)

    def close(self):
        if self._h:
            for p in list(self._plans):
                p.close()
            lib.bb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- post-inference ---------------------------------------------------------------
    def post_run(self, d_scores: int, B: int, Cc: int, valid_B: int, cfg: "PostConfig",
                 d_mask: Optional[int] = None, d_species_keep: Optional[int] = None):
        """Host-output variant: returns (index [valid,k] u32, conf [valid,k] f32, count [valid] u32)."""
        k = cfg.top_k
        idx = np.zeros((max(valid_B, 1), k), np.uint32)
        conf = np.zeros((max(valid_B, 1), k), np.float32)
        cnt = np.zeros(max(valid_B, 1), np.uint32)
        c = cfg.to_c()
        check(lib.bb_post_run(self._h, C.c_void_p(d_scores), B, Cc, valid_B, C.byref(c),
                              C.c_void_p(d_mask) if d_mask else None, C.c_void_p(d_species_keep) if d_species_keep else None,
                              idx.ctypes.data_as(_lib.u32p), conf.ctypes.data_as(_lib.f32p), cnt.ctypes.data_as(_lib.u32p)),
              self._h)
        return idx[:valid_B], conf[:valid_B], cnt[:valid_B]

    def dense_run(self, d_x: int, B: int, K: int, d_W: int, d_b: Optional[int], N: int, activation: int, d_out: int):
        """out[B,N] = act(x[B,K] W[K,N] + b[N]) on the device (geomodel forward, bat head)."""
        check(lib.bb_dense_run(self._h, C.c_void_p(d_x), B, K, C.c_void_p(d_W), C.c_void_p(d_b) if d_b else None, N,
                               activation, C.c_void_p(d_out)), self._h)

    def post_run_device(self, d_scores: int, B: int, Cc: int, valid_B: int, cfg: "PostConfig",
                        d_mask: Optional[int], d_species_keep: Optional[int], d_index: int, d_conf: int, d_count: int):
        c = cfg.to_c()
        check(lib.bb_post_run_device(self._h, C.c_void_p(d_scores), B, Cc, valid_B, C.byref(c),
                                     C.c_void_p(d_mask) if d_mask else None,
                                     C.c_void_p(d_species_keep) if d_species_keep else None,
                                     C.c_void_p(d_index), C.c_void_p(d_conf), C.c_void_p(d_count)), self._h)


@dataclass
class PostConfig:
    activation: int = ACT_SIGMOID
    min_confidence: float = 0.1          # src/constants.rs:25
    top_k: int = 5                       # src/constants.rs:178
    range_threshold: float = 0.01        # src/constants.rs:333
    keep_unmatched: bool = True
    rerank: bool = False

    def to_c(self) -> PostCfg:
        return PostCfg(self.activation, self.min_confidence, self.top_k, self.range_threshold,
                       int(self.keep_unmatched), int(self.rerank))


@dataclass
class Segments:
    """Result of one front-end run: a device tensor [rows, segment_samples] f32 plus host tables."""
    device_ptr: int
    nseg: int
    rows: int
    segment_samples: int
    start_sample: np.ndarray
    start_time: np.ndarray
    end_time: np.ndarray
    consumed_frames: int
    device: int = 0

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.rows, self.segment_samples), "typestr": "<f4", "data": (self.device_ptr, False),
                "version": 3, "strides": None}

    def torch(self):
        import torch
        return torch.as_tensor(self, device=f"cuda:{self.device}")


class FrontEndPlan:
    """``bb_plan``: one (src_rate, channels, fmt, tgt_rate, segment, overlap) configuration."""

    def __init__(self, ctx: Context, src_rate: int, channels: int, fmt: int, tgt_rate: int,
                 segment_samples: int, overlap_samples: int):
        self.ctx = ctx
        self._h = C.c_void_p()
        check(lib.bb_plan_create(ctx.handle, src_rate, channels, fmt, tgt_rate, segment_samples, overlap_samples,
                                 C.byref(self._h)), ctx.handle)
        ctx._plans.add(self)
        self.src_rate, self.tgt_rate, self.channels, self.fmt = src_rate, tgt_rate, channels, fmt
        self.segment_samples, self.overlap_samples = segment_samples, overlap_samples
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.bb_plan_source_window(self._h, C.byref(a), C.byref(b)))
        self.src_segment, self.src_overlap = a.value, b.value

    def segment_count(self, total_frames: int) -> int:
        n = C.c_uint64()
        check(lib.bb_plan_segment_count(self._h, total_frames, C.byref(n)))
        return n.value

    def describe(self) -> str:
        """Which kernel, transform sizes and blocking this plan launches (``bb_plan_describe``)."""
        buf = C.create_string_buffer(320)
        check(lib.bb_plan_describe(self._h, buf, 320))
        return buf.value.decode()

    def run(self, pcm, frames: Optional[int] = None, *, is_device: Optional[bool] = None, first_start_sample: int = 0,
            is_eof: bool = True, pad_to_batch: int = 0, out_ptr: Optional[int] = None,
            out_capacity_rows: int = 0, want_tables: bool = True) -> Segments:
        """``pcm``: numpy array (host, interleaved) or anything with ``data_ptr()`` (torch CUDA tensor) or an int
        device pointer (then ``frames`` and ``is_device=True`` are required)."""
        keep = None
        if isinstance(pcm, np.ndarray):
            if _NP_FMT.get(pcm.dtype) != self.fmt:
                raise BirdaError(-4, f"pcm dtype {pcm.dtype} does not match the plan's sample format")
            pcm = np.ascontiguousarray(pcm)
            keep = pcm
            ptr = pcm.ctypes.data
            n = pcm.size // (self.channels * (3 if self.fmt == FMT_S24 else 1)) if frames is None else frames
            dev = False if is_device is None else is_device
        elif hasattr(pcm, "data_ptr"):
            ptr = pcm.data_ptr()
            n = pcm.numel() // (self.channels * (3 if self.fmt == FMT_S24 else 1)) if frames is None else frames
            dev = bool(pcm.is_cuda) if is_device is None else is_device
            keep = pcm
        else:
            ptr, n, dev = int(pcm), int(frames), bool(is_device)
        nseg_max = rules.segment_count(n, self.src_segment, self.src_overlap)
        rows_max = nseg_max
        if pad_to_batch > 1 and nseg_max % pad_to_batch:
            rows_max = (nseg_max // pad_to_batch + 1) * pad_to_batch
        cap = max(rows_max, out_capacity_rows, 1)
        ss = np.zeros(cap, np.uint64)
        st = np.zeros(cap, np.float32)
        et = np.zeros(cap, np.float32)
        d_seg = C.c_void_p()
        nseg, rows, consumed = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.bb_frontend_run(self._h, C.c_void_p(ptr), n, int(dev), first_start_sample, int(is_eof), pad_to_batch,
                                  C.c_void_p(out_ptr) if out_ptr else None, cap,
                                  C.byref(d_seg), ss.ctypes.data_as(_lib.u64p), st.ctypes.data_as(_lib.f32p),
                                  et.ctypes.data_as(_lib.f32p), C.byref(nseg), C.byref(rows), C.byref(consumed)),
              self.ctx.handle)
        self._keepalive = keep          # host PCM must outlive the async H2D copy
        return Segments(d_seg.value or 0, nseg.value, rows.value, self.segment_samples, ss[: nseg.value],
                        st[: nseg.value], et[: nseg.value], consumed.value, self.ctx.device)

    def close(self):
        if self._h:
            lib.bb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MelSpec:
    """Spectrogram prefix on the device (``bb_melspec``): framed STFT power + tcgen05 mel projection.
    ``window``: [n_fft] f32, ``mel_weights``: [n_mels, n_fft/2 + 1] f32 (host arrays, copied)."""

    def __init__(self, ctx: Context, n_fft: int, hop: int, n_frames: int, window: np.ndarray, mel_weights: np.ndarray,
                 power: float = 2.0, log_mode: int = 0, log_eps: float = 1e-6):
        self.ctx = ctx
        w = np.ascontiguousarray(window, dtype=np.float32)
        mw = np.ascontiguousarray(mel_weights, dtype=np.float32)
        if w.shape != (n_fft,) or mw.ndim != 2 or mw.shape[1] != n_fft // 2 + 1:
            raise ValueError("window must be [n_fft] and mel_weights [n_mels, n_fft/2 + 1]")
        self.n_fft, self.hop, self.n_frames, self.n_mels = n_fft, hop, n_frames, int(mw.shape[0])
        cfg = MelSpecCfg(n_fft, hop, n_frames, self.n_mels, power, log_mode, log_eps)
        self._h = C.c_void_p()
        check(lib.bb_melspec_create(ctx.handle, C.byref(cfg), w.ctypes.data_as(_lib.f32p), mw.ctypes.data_as(_lib.f32p),
                                    C.byref(self._h)), ctx.handle)
        ctx._plans.add(self)

    def info(self) -> Tuple[int, int, int]:
        """(first bin with mel weight, bins with mel weight, GEMM K after padding to 32)."""
        a, b_, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib.bb_melspec_info(self._h, C.byref(a), C.byref(b_), C.byref(c)), self.ctx.handle)
        return a.value, b_.value, c.value

    def run(self, d_segments: int, rows: int, samples: int, d_out: int) -> None:
        """d_segments: device [rows, samples] f32; d_out: device [rows, n_mels, n_frames] f32.  Asynchronous."""
        check(lib.bb_melspec_run(self._h, C.c_void_p(d_segments), rows, samples, C.c_void_p(d_out)), self.ctx.handle)

    def close(self):
        if self._h:
            lib.bb_melspec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
