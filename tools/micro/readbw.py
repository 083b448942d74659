#!/usr/bin/env python3
"""Page cache -> host buffer read bandwidth of bb_wav_read_parallel on this box (pinned and pageable destinations)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import birda_b200 as b
from birda_b200 import _lib
from birda_b200.synth import synth_pcm, write_wav

path = "/dev/shm/readbw.wav"
pcm = np.tile(synth_pcm(1, 30.0, 96_000, 2), 20)      # 230 MB
write_wav(path, pcm, 96_000, 2)
info = b.wav_probe(path)
nbytes = info.frames * 4
pinned = C.c_void_p()
_lib.check(_lib.lib.bb_host_alloc(nbytes, C.byref(pinned)))
pageable = np.zeros(nbytes, np.uint8)
for name, dst in (("pinned", pinned), ("pageable", C.c_void_p(pageable.ctypes.data))):
    for th in (1, 2, 4, 8, 16):
        _lib.lib.bb_wav_read_parallel(path.encode(), C.byref(info), 0, info.frames, dst, th)
        t0 = time.perf_counter()
        for _ in range(5):
            _lib.lib.bb_wav_read_parallel(path.encode(), C.byref(info), 0, info.frames, dst, th)
        dt = (time.perf_counter() - t0) / 5
        print(f"{name:9s} threads {th:2d}: {nbytes / dt / 1e9:6.1f} GB/s")
t0 = time.perf_counter()
for _ in range(5):
    np.copyto(pageable, pageable[::-1][::-1] if False else pageable)
a = np.zeros(nbytes, np.uint8)
t0 = time.perf_counter()
for _ in range(5):
    np.copyto(a, pageable)
print(f"numpy memcpy 1 thread: {nbytes / ((time.perf_counter() - t0) / 5) / 1e9:.1f} GB/s")
os.remove(path)
