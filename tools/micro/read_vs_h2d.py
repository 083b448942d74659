#!/usr/bin/env python3
"""Do page-cache reads and pinned->device copies overlap on this box?  Each alone, then both at once."""
import ctypes as C
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import birda_b200 as b
from birda_b200 import _lib
from birda_b200.synth import synth_pcm, write_wav

path = "/dev/shm/readbw.wav"
write_wav(path, np.tile(synth_pcm(1, 30.0, 96_000, 2), 20), 96_000, 2)
info = b.wav_probe(path)
nbytes = info.frames * 4
pinned = C.c_void_p(); _lib.check(_lib.lib.bb_host_alloc(nbytes, C.byref(pinned)))
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
TH = int(sys.argv[1]) if len(sys.argv) > 1 else 12

def reads(n, out):
    t0 = time.perf_counter()
    for _ in range(n):
        _lib.lib.bb_wav_read_parallel(path.encode(), C.byref(info), 0, info.frames, pinned, TH)
    out.append(n * nbytes / (time.perf_counter() - t0) / 1e9)

def copies(n, out):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    out.append(n * nbytes / (time.perf_counter() - t0) / 1e9)

r, c = [], []
reads(10, r); copies(20, c)
print(f"alone: read ({TH} threads) {r[0]:.1f} GB/s, H2D {c[0]:.1f} GB/s")
r, c = [], []
t = threading.Thread(target=reads, args=(20, r)); t.start(); copies(40, c); t.join()
print(f"together: read {r[0]:.1f} GB/s, H2D {c[0]:.1f} GB/s")
os.remove(path)
