#!/usr/bin/env python3
"""BASELINE config 5 in miniature through the native C++ pipeline (bb_pipeline_*): 10-minute files cycling over six
sample rates and mono / stereo, pinned host PCM, a stand-in classifier that returns resident scores (the model is
not part of this path).  Prints per-file wall time and audio-hours/sec.
usage: python tools/prof_c5.py [files] [minutes_per_file]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import birda_b200 as b
from birda_b200.pipeline import NativePipeline, ProcessingConfig

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 24
minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
C = 6522
ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
scores = torch.randn((64, C), device="cuda") * 2 - 6


def classifier(x):                      # [rows, 144000] -> [rows, C]; resident scores: the forward is not on this path
    return scores[: x.shape[0]]


mask = torch.rand(C, device="cuda") ** 2
cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=0.0, batch_size=64, min_confidence=0.1,
                       d_mask=mask.data_ptr(), range_threshold=0.01)
kinds = [(16_000, 1), (22_050, 2), (32_000, 1), (44_100, 2), (48_000, 1), (96_000, 2)]
files = []
for i in range(nfiles):
    sr, ch = kinds[i % len(kinds)]
    n = int(minutes * 60 * sr) * ch
    t = torch.empty(n, dtype=torch.int16).pin_memory()
    t.copy_((torch.randn(n) * 3000).to(torch.int16))
    files.append((sr, ch, t.numpy()))
nat = NativePipeline(ctx, cfg, classifier)
for sr, ch, pcm in files[: len(kinds)]:           # first pass builds the plans
    nat.process_pcm(pcm, ch, sr, b.FMT_S16)
torch.cuda.synchronize()
t0 = time.perf_counter()
segs = 0
per_kind = {}
for sr, ch, pcm in files:
    t1 = time.perf_counter()
    r = nat.process_pcm(pcm, ch, sr, b.FMT_S16)
    per_kind.setdefault((sr, ch), []).append(time.perf_counter() - t1)
    segs += r.segments
dt = time.perf_counter() - t0
hours = nfiles * minutes / 60.0
print(f"c5 mini: {nfiles} files x {minutes} min, {segs} windows, {nat.plans_created} plans: {dt * 1e3 / nfiles:.2f} ms per file, "
      f"{hours / dt:.1f} audio-h/s (1 GPU, pinned PCM, stand-in classifier)")
for k, v in sorted(per_kind.items()):
    mb = minutes * 60 * k[0] * k[1] * 2 / 1e6
    print(f"  {k[0]:6d} Hz x{k[1]}: {np.mean(v) * 1e3:7.2f} ms per file ({mb:6.1f} MB PCM -> {mb / np.mean(v) / 1e3:5.1f} GB/s)")
nat.close(); ctx.close()
