"""Host-side mirror of the reference's per-file pipeline over the C ABI (tests, bench, Python hosts).

Follows `process_file` / `run_streaming_inference` / `process_batch`
(src/pipeline/processor.rs:418-796, :114-190, :220-410): estimate -> effective batch size -> front end
-> batches of `batch_size` rows (last one padded with silence) -> classifier -> post step on the valid
rows -> detections sorted by (start_time asc, confidence desc).  The classifier is any callable mapping
a device tensor [B, samples] to scores [B, C] (ONNX Runtime with IoBinding in the reference's world; a
stand-in in the tests) — it is not part of this library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .api import ACT_SIGMOID, Context, FrontEndPlan, PostConfig, rules


@dataclass
class Detection:
    """src/output/types.rs:8-23 (fields the writers use)."""
    scientific_name: str
    common_name: str
    confidence: float
    start_time: float
    end_time: float
    index: int
    segment: int


def split_label(label: str) -> Tuple[str, str]:
    """Detection::from_label: split at the first '_' (src/output/types.rs:58-79)."""
    if "_" in label:
        a, b = label.split("_", 1)
        return a, b
    return label, label


@dataclass
class ProcessingConfig:
    """The hot-path subset of src/pipeline/config.rs:32-66 plus the classifier's model facts."""
    target_rate: int = 48_000
    segment_duration: float = 3.0
    overlap: float = 0.0                       # src/constants.rs:28
    batch_size: int = 64
    min_confidence: float = 0.1                # src/constants.rs:25
    top_k: int = 5                             # src/constants.rs:178
    activation: int = ACT_SIGMOID
    bat_mode: bool = False                     # processor.rs:464-475
    range_threshold: float = 0.01
    keep_unmatched: bool = True
    rerank: bool = False
    labels: Optional[Sequence[str]] = None
    d_mask: Optional[int] = None               # device pointer, [C] f32, NaN = no geomodel entry
    d_species_keep: Optional[int] = None       # device pointer, [C] u8


@dataclass
class ProcessResult:
    detections: List[Detection] = field(default_factory=list)
    segments: int = 0
    batches: int = 0
    effective_batch_size: int = 0


def sort_detections(dets: List[Detection]) -> List[Detection]:
    """src/pipeline/processor.rs:178-187 (stable here; the reference's sort is unstable)."""
    return sorted(dets, key=lambda d: (d.start_time, -d.confidence))


class FilePipeline:
    """One GPU's worker: owns a Context; plans are cached per (rate, channels, format)."""

    def __init__(self, ctx: Context, cfg: ProcessingConfig, classifier: Callable):
        self.ctx, self.cfg, self.classifier = ctx, cfg, classifier
        self._plans = {}

    def _plan(self, source_rate: int, channels: int, fmt: int) -> FrontEndPlan:
        cfg = self.cfg
        target = source_rate if cfg.bat_mode else cfg.target_rate
        seg, ovl = rules.segment_samples(cfg.segment_duration, cfg.overlap, target, cfg.bat_mode)
        key = (source_rate, channels, fmt, target, seg, ovl)
        if key not in self._plans:
            self._plans[key] = FrontEndPlan(self.ctx, source_rate, channels, fmt, target, seg, ovl)
        return self._plans[key]

    def process_pcm(self, pcm, channels: int, source_rate: int, fmt: int, frames: Optional[int] = None) -> ProcessResult:
        cfg = self.cfg
        plan = self._plan(source_rate, channels, fmt)
        n_frames = (pcm.size if isinstance(pcm, np.ndarray) else pcm.numel()) // channels if frames is None else frames
        seg_dur = float(rules_bat_duration()) if cfg.bat_mode else cfg.segment_duration
        est = rules.estimate_segment_count(n_frames / source_rate, seg_dur, cfg.overlap)        # processor.rs:525
        B = max(rules.effective_batch_size(cfg.batch_size, est), 1)                              # processor.rs:531-545
        segs = plan.run(pcm, frames=n_frames, pad_to_batch=B)
        res = ProcessResult(segments=segs.nseg, effective_batch_size=B)
        if segs.nseg == 0:
            return res
        # the classifier runs on torch's current stream: when the context owns a different (non-blocking) stream the
        # two must be ordered by hand — packed windows finished before the classifier reads them, scores finished
        # before the post kernel reads them (a Rust host with ORT on the ctx stream needs neither)
        differs = ctx_stream_differs(self.ctx)
        if differs:
            self.ctx.sync()
        x = segs.torch()
        post = PostConfig(cfg.activation, cfg.min_confidence, cfg.top_k, cfg.range_threshold, cfg.keep_unmatched, cfg.rerank)
        dets: List[Detection] = []
        for first in range(0, segs.nseg, B):
            valid = min(B, segs.nseg - first)
            scores = self.classifier(x[first:first + B])                                        # [B, C] on the device
            if differs:
                import torch
                torch.cuda.current_stream().synchronize()
            Bc, C = int(scores.shape[0]), int(scores.shape[1])
            idx, conf, cnt = self.ctx.post_run(scores.data_ptr(), Bc, C, valid, post, cfg.d_mask, cfg.d_species_keep)
            res.batches += 1
            for r in range(valid):                                                               # processor.rs:363-385
                s = first + r
                for j in range(int(cnt[r])):
                    c = float(conf[r, j])
                    if c >= cfg.min_confidence:
                        i = int(idx[r, j])
                        sci, com = split_label(cfg.labels[i]) if cfg.labels is not None else (str(i), str(i))
                        dets.append(Detection(sci, com, c, float(segs.start_time[s]), float(segs.end_time[s]), i, s))
        res.detections = sort_detections(dets)
        return res

    def close(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()


def rules_bat_duration() -> np.float32:
    """bat::SEGMENT_DURATION = 144000 / 256000 as f32 (src/constants.rs:525-535)."""
    return np.float32(144_000) / np.float32(256_000)


class _DevView:
    """__cuda_array_interface__ view of a device pointer (for torch.as_tensor)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class NativePipeline:
    """The C++ per-file pipeline of the library (csrc/pipeline.cpp: bb_pipeline_*) with a Python classifier
    callback.  Same contract as FilePipeline; detections come back as Detection objects."""

    def __init__(self, ctx: Context, cfg: ProcessingConfig, classifier: Callable):
        import ctypes as C

        from . import _lib
        self.ctx, self.cfg, self.classifier = ctx, cfg, classifier
        self._scores = None
        self._error = None

        def _cb(user, d_segments, rows, samples, d_scores, classes):
            try:
                import torch
                differs = ctx_stream_differs(ctx)
                if differs:                            # the packed windows are written on the ctx stream
                    ctx.sync()
                x = torch.as_tensor(_DevView(d_segments, (rows, samples)), device=f"cuda:{ctx.device}")
                s = self.classifier(x).contiguous().float()
                self._scores = s                       # keep alive until the next call
                d_scores[0] = s.data_ptr()
                classes[0] = int(s.shape[1])
                if differs:                            # ... and the post kernel reads the scores on it
                    torch.cuda.current_stream().synchronize()
                return 0
            except Exception as e:                      # never let an exception cross the C boundary
                self._error = e
                return 1

        from .api import StandIn
        native = isinstance(classifier, StandIn)           # a bb_classify_fn of the library: no Python in the loop
        self._cb = classifier.fn if native else _lib.CLASSIFY_FN(_cb)
        c = _lib.PipelineCfg(cfg.target_rate, cfg.segment_duration, cfg.overlap, cfg.batch_size, int(cfg.bat_mode),
                             PostConfig(cfg.activation, cfg.min_confidence, cfg.top_k, cfg.range_threshold,
                                        cfg.keep_unmatched, cfg.rerank).to_c(), cfg.d_mask, cfg.d_species_keep)
        self._h = C.c_void_p()
        _lib.check(_lib.lib.bb_pipeline_create(ctx.handle, C.byref(c), self._cb, classifier.handle if native else None,
                                               C.byref(self._h)), ctx.handle)

    def _collect(self, call) -> ProcessResult:
        import ctypes as C

        from . import _lib
        cap = 1 << 16
        while True:
            buf = (_lib.DetectionC * cap)()
            nd, ns, bu = C.c_uint64(), C.c_uint64(), C.c_uint32()
            rc = call(buf, cap, C.byref(nd), C.byref(ns), C.byref(bu))
            if rc == -9 and nd.value > cap:             # BB_ERR_CAPACITY: retry with the reported size
                cap = int(nd.value)
                continue
            if rc != 0:
                msg = _lib.lib.bb_pipeline_last_error(self._h).decode("utf-8", "replace")
                if self._error is not None:
                    raise self._error
                raise _lib.BirdaError(rc, msg)
            break
        res = ProcessResult(segments=int(ns.value), effective_batch_size=int(bu.value))
        labels = self.cfg.labels
        for i in range(int(nd.value)):
            d = buf[i]
            sci, com = split_label(labels[d.index]) if labels is not None else (str(d.index), str(d.index))
            res.detections.append(Detection(sci, com, float(d.confidence), float(d.start_time), float(d.end_time), int(d.index), int(d.segment)))
        res.batches = -(-res.segments // max(res.effective_batch_size, 1))
        return res

    def set_batch_hooks(self, before: Optional[Callable] = None, after: Optional[Callable] = None) -> None:
        """``before(batch_rows, valid_rows, first_segment)`` / ``after(...)`` around every inference batch — the seam
        where the reference arms and drops its watchdog (src/pipeline/processor.rs:263-277)."""
        import ctypes as C

        from . import _lib
        mk = lambda f: _lib.BATCH_HOOK(lambda user, rows, valid, first: f(int(rows), int(valid), int(first))) if f else C.cast(None, _lib.BATCH_HOOK)
        self._hooks = (mk(before), mk(after))
        _lib.lib.bb_pipeline_set_batch_hooks(self._h, self._hooks[0], self._hooks[1], None)

    def set_batch_timeout(self, timeout_ms: int, on_fire: Optional[Callable] = None) -> None:
        """Library-kept watchdog per batch.  ``on_fire(timeout_secs, batch)``; None = the reference's exit(1)."""
        import ctypes as C

        from . import _lib
        self._on_fire = _lib.WATCHDOG_FN(lambda user, secs, batch: on_fire(int(secs), int(batch))) if on_fire else C.cast(None, _lib.WATCHDOG_FN)
        _lib.lib.bb_pipeline_set_batch_timeout(self._h, timeout_ms, self._on_fire, None)

    def process_pcm(self, pcm: np.ndarray, channels: int, source_rate: int, fmt: int) -> ProcessResult:
        import ctypes as C

        from . import _lib
        pcm = np.ascontiguousarray(pcm)
        frames = pcm.size // (channels * (3 if fmt == 4 else 1))
        return self._collect(lambda buf, cap, nd, ns, bu: _lib.lib.bb_pipeline_process_pcm(
            self._h, C.c_void_p(pcm.ctypes.data), frames, source_rate, channels, fmt, buf, cap, nd, ns, bu))

    def process_wav(self, path: str, piece_frames: int = 0) -> ProcessResult:
        from . import _lib
        return self._collect(lambda buf, cap, nd, ns, bu: _lib.lib.bb_pipeline_process_wav(
            self._h, path.encode(), piece_frames, buf, cap, nd, ns, bu))

    @property
    def plans_created(self) -> int:
        """Front-end plans built so far (they are cached per source rate / channels / format)."""
        from . import _lib
        return int(_lib.lib.bb_pipeline_plans_created(self._h))

    def close(self):
        from . import _lib
        if self._h:
            _lib.lib.bb_pipeline_destroy(self._h)
            self._h = None


class NativePool:
    """The C++ multi-GPU directory runner (csrc/pool.cpp: bb_pool_*): one worker thread + context + pipeline per entry of
    ``devices``, WAV files handed out longest first.  ``classifiers[i]`` serves ``devices[i]`` and is called from that
    worker's thread (after the context's stream was synchronised); it must return scores whose computation is finished."""

    def __init__(self, devices: List[int], cfgs: List[ProcessingConfig], classifiers: List[Callable]):
        import ctypes as C

        from . import _lib
        assert len(devices) == len(cfgs) == len(classifiers) and devices
        self.devices, self.cfgs, self.classifiers = list(devices), list(cfgs), list(classifiers)
        self._scores = [None] * len(devices)
        self._errors: list = []

        def _cb(user, d_segments, rows, samples, d_scores, classes):
            try:
                import torch
                i = int(user or 0) - 1                 # users[i] = i + 1 (a null user pointer would read as None)
                dev = self.devices[i]
                with torch.cuda.device(dev):
                    x = torch.as_tensor(_DevView(d_segments, (rows, samples)), device=f"cuda:{dev}")
                    s = self.classifiers[i](x).contiguous().float()
                    torch.cuda.current_stream().synchronize()
                self._scores[i] = s                    # keep alive until this worker's next call
                d_scores[0] = s.data_ptr()
                classes[0] = int(s.shape[1])
                return 0
            except Exception as e:                      # never let an exception cross the C boundary
                self._errors.append(e)
                return 1

        from .api import StandIn
        native = all(isinstance(c, StandIn) for c in classifiers)      # bb_classify_fn of the library: no Python in the loop
        self._cb = classifiers[0].fn if native else _lib.CLASSIFY_FN(_cb)
        n = len(devices)
        c_cfgs = (_lib.PipelineCfg * n)(*[
            _lib.PipelineCfg(c.target_rate, c.segment_duration, c.overlap, c.batch_size, int(c.bat_mode),
                             PostConfig(c.activation, c.min_confidence, c.top_k, c.range_threshold, c.keep_unmatched, c.rerank).to_c(),
                             c.d_mask, c.d_species_keep) for c in cfgs])
        users = (C.c_void_p * n)(*[(classifiers[i].handle if native else C.c_void_p(i + 1)) for i in range(n)])
        devs = (C.c_int32 * n)(*devices)
        self._h = C.c_void_p()
        _lib.check(_lib.lib.bb_pool_create(devs, n, c_cfgs, self._cb, users, C.byref(self._h)))

    @property
    def kernel_launches(self) -> int:
        from . import _lib
        return int(_lib.lib.bb_pool_kernel_launches(self._h))

    def worker_stream(self, i: int) -> int:
        """cudaStream_t (as an int) of worker ``i``'s context."""
        from . import _lib
        return int(_lib.lib.bb_ctx_stream(_lib.lib.bb_pool_worker_ctx(self._h, i)) or 0)

    def bind_standins_to_worker_streams(self) -> None:
        """Native stand-in classifiers queue on their worker's stream; the pool stops waiting around every batch."""
        from . import _lib
        from .api import StandIn
        assert all(isinstance(c, StandIn) for c in self.classifiers)
        for i, c in enumerate(self.classifiers):
            _lib.lib.bb_standin_use_stream(c.handle, self.worker_stream(i), 1)
        _lib.lib.bb_pool_set_stream_ordered(self._h, 1)

    def process_wavs_counts(self, paths: List[str]):
        """The same run without turning detections into Python objects: [(segments, detections, device)] per file
        (benches time the library, not the conversion)."""
        import ctypes as C

        from . import _lib
        n = len(paths)
        arr = (C.c_char_p * n)(*[p.encode() for p in paths])
        res = (_lib.PoolResult * n)()
        rc = _lib.lib.bb_pool_process_wavs(self._h, arr, n, res)
        try:
            for i in range(n):
                if res[i].status != 0:
                    raise _lib.BirdaError(int(res[i].status), f"{paths[i]}: {res[i].error.decode('utf-8', 'replace')}")
            out = [(int(res[i].n_segments), int(res[i].n_detections), int(res[i].device)) for i in range(n)]
        finally:
            _lib.lib.bb_pool_free_results(res, n)
        assert rc == 0
        return out

    def process_wavs(self, paths: List[str]) -> List[ProcessResult]:
        import ctypes as C

        from . import _lib
        n = len(paths)
        arr = (C.c_char_p * n)(*[p.encode() for p in paths])
        res = (_lib.PoolResult * n)()
        rc = _lib.lib.bb_pool_process_wavs(self._h, arr, n, res)
        out = []
        try:
            for i in range(n):
                r = res[i]
                if r.status != 0:
                    if self._errors:
                        raise self._errors[0]
                    raise _lib.BirdaError(int(r.status), f"{paths[i]}: {r.error.decode('utf-8', 'replace')}")
                pr = ProcessResult(segments=int(r.n_segments), effective_batch_size=int(r.batch_used))
                pr.device = int(r.device)
                labels = self.cfgs[0].labels
                for j in range(int(r.n_detections)):
                    d = r.detections[j]
                    sci, com = split_label(labels[d.index]) if labels is not None else (str(d.index), str(d.index))
                    pr.detections.append(Detection(sci, com, float(d.confidence), float(d.start_time), float(d.end_time), int(d.index), int(d.segment)))
                pr.batches = -(-pr.segments // max(pr.effective_batch_size, 1))
                out.append(pr)
        finally:
            _lib.lib.bb_pool_free_results(res, n)
        assert rc == 0
        return out

    def close(self):
        from . import _lib
        if self._h:
            _lib.lib.bb_pool_destroy(self._h)
            self._h = None


def ctx_stream_differs(ctx: Context) -> bool:
    """True when the context runs on its own stream: the classifier's torch work (current stream) must be
    finished before the library's post kernel reads the scores."""
    import torch

    from . import _lib
    return int(_lib.lib.bb_ctx_stream(ctx.handle) or 0) != int(torch.cuda.current_stream().cuda_stream)
