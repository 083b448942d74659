"""ctypes access to the C restatement (oracle/c/oracle_birda.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle_birda.so")
_lib = None

_FMT = {np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.float32): 3}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        subprocess.run(["make"], cwd=_DIR, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    lib = C.CDLL(_SO)
    lib.orc_frontend.restype = C.c_int64
    lib.orc_frontend.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64,
                                 C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_resample.restype = C.c_int64
    lib.orc_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
    lib.orc_resampler_info.restype = C.c_int32
    lib.orc_resampler_info.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float), C.c_void_p, C.c_int32]
    lib.orc_post.restype = C.c_int32
    lib.orc_post.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p,
                             C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    _lib = lib
    return lib


def frontend(pcm: np.ndarray, channels: int, sr: int, tr: int, seg: int, ovl: int, threads: int = 1,
             compute: bool = True):
    """(segments [nseg, seg] f32 or None, start_sample u64, start_time f32, end_time f32)"""
    lib = load()
    pcm = np.ascontiguousarray(pcm)
    frames = pcm.size // channels
    n = lib.orc_frontend(pcm.ctypes.data, _FMT[pcm.dtype], channels, frames, sr, tr, seg, ovl, None, 0, None, None, None, 1)
    if n < 0:
        raise ValueError(f"orc_frontend error {n}")
    ss = np.zeros(max(n, 1), np.uint64); st = np.zeros(max(n, 1), np.float32); et = np.zeros(max(n, 1), np.float32)
    out = np.zeros((max(n, 1), seg), np.float32) if compute else None
    lib.orc_frontend(pcm.ctypes.data, _FMT[pcm.dtype], channels, frames, sr, tr, seg, ovl,
                     out.ctypes.data if compute else None, n, ss.ctypes.data, st.ctypes.data, et.ctypes.data, threads)
    return (out[:n] if compute else None), ss[:n], st[:n], et[:n]


def resample(x: np.ndarray, sr: int, tr: int) -> np.ndarray:
    lib = load()
    x = np.ascontiguousarray(x, np.float32)
    cap = int(x.size * tr / sr) + 70000
    out = np.zeros(cap, np.float32)
    n = lib.orc_resample(x.ctypes.data, x.size, sr, tr, out.ctypes.data, cap)
    return out[:n].copy()


def resampler_info(sr: int, tr: int):
    lib = load()
    a, b, c = C.c_int32(), C.c_int32(), C.c_float()
    if lib.orc_resampler_info(sr, tr, C.byref(a), C.byref(b), C.byref(c), None, 0) != 0:
        raise ValueError("unsupported rates")
    taps = np.zeros(a.value, np.float32)
    lib.orc_resampler_info(sr, tr, C.byref(a), C.byref(b), C.byref(c), taps.ctypes.data, a.value)
    return a.value, b.value, np.float32(c.value), taps


def post(scores: np.ndarray, valid: int, act: int, min_conf: float, top_k: int = 5, mask: Optional[np.ndarray] = None,
         keep: Optional[np.ndarray] = None, threshold: float = 0.01, keep_unmatched: bool = True, rerank: bool = False,
         threads: int = 1):
    lib = load()
    scores = np.ascontiguousarray(scores, np.float32)
    idx = np.zeros((max(valid, 1), top_k), np.uint32); conf = np.zeros((max(valid, 1), top_k), np.float32)
    cnt = np.zeros(max(valid, 1), np.uint32)
    m = np.ascontiguousarray(mask, np.float32) if mask is not None else None
    k = np.ascontiguousarray(keep, np.uint8) if keep is not None else None
    rc = lib.orc_post(scores.ctypes.data, scores.shape[1], valid, act, min_conf, top_k,
                      m.ctypes.data if m is not None else None, k.ctypes.data if k is not None else None,
                      threshold, int(keep_unmatched), int(rerank), idx.ctypes.data, conf.ctypes.data, cnt.ctypes.data, threads)
    if rc != 0:
        raise ValueError("orc_post failed")
    return idx[:valid], conf[:valid], cnt[:valid]
