#!/usr/bin/env python3
"""Driver for timing / ncu captures of K5 (spectrogram prefix) on a C2-shaped batch of packed windows.
usage: python tools/prof_mel.py [rows] [iters] [low|high|full]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import birda_b200 as b

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
band = sys.argv[3] if len(sys.argv) > 3 else "low"
n_fft, hop, n_frames, n_mels, fmin, fmax = {"low": (2048, 278, 511, 96, 0.0, 3000.0), "high": (1024, 280, 511, 96, 500.0, 15000.0),
                                            "full": (2048, 278, 511, 128, 0.0, 24000.0)}[band]
rate, samples = 48_000, 144_000


def hz_to_mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


bins = n_fft // 2 + 1
m = hz_to_mel(np.linspace(0.0, rate / 2.0, bins))
edges = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
mw = np.zeros((n_mels, bins), dtype=np.float32)
for i in range(n_mels):
    mw[i] = np.maximum(0.0, np.minimum((m - edges[i]) / (edges[i + 1] - edges[i]), (edges[i + 2] - m) / (edges[i + 2] - edges[i + 1])))
mw[:, 0] = 0
win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
x = torch.randn((rows, samples), device="cuda") * 0.1
out = torch.empty((rows, n_mels, n_frames), device="cuda")
ms = b.MelSpec(ctx, n_fft, hop, n_frames, win, mw)
lo, nb, kpad = ms.info()
for _ in range(2):
    ms.run(x.data_ptr(), rows, samples, out.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ms.run(x.data_ptr(), rows, samples, out.data_ptr())
e1.record(); torch.cuda.synchronize()
ms_it = e0.elapsed_time(e1) / iters
frames = rows * n_frames
gemm_flops = 3 * 2.0 * frames * kpad * n_mels
print(f"mel {band}: {rows} rows x {n_frames} frames, n_fft {n_fft}, bins [{lo}, {lo + nb}) -> K {kpad}, {n_mels} mels: "
      f"{ms_it:.3f} ms/iter ({rows * 1.5 / 3600 / (ms_it / 1e3):.0f} audio-h/s at hop 1.5 s); GEMM {gemm_flops / 1e9:.1f} GFLOP (3 tf32 products)")
