#!/usr/bin/env python3
"""Small driver for ncu captures: runs each kernel of the path a few times on a C2-shaped sample.
usage: python tools/prof_run.py [minutes_of_audio] [iters] [config]   config: c2 | c3 | c4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import birda_b200 as b

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
config = sys.argv[3] if len(sys.argv) > 3 else "c2"
dev = torch.device("cuda", 0)
ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
if config == "c2":
    sr, ch, tr, seg, ovl, C = 44_100, 2, 48_000, 144_000, 72_000, 6522
elif config == "c3":
    sr, ch, tr, seg, ovl, C = 48_000, 1, 32_000, 160_000, 0, 14795
elif config.startswith("r"):          # r96000 / r32000 / r16000 / r22050: mono s16 at that rate -> 48 kHz BirdNET windows, overlap 0
    sr, ch, tr, seg, ovl, C = int(config[1:]), 1, 48_000, 144_000, 0, 6522
else:
    sr, ch, tr, seg, ovl, C = 256_000, 1, 256_000, 144_000, 36_000, 6522
n = int(minutes * 60 * sr)
pcm = (torch.randn(n * ch, device=dev) * 3000).to(torch.int16)
plan = b.FrontEndPlan(ctx, sr, ch, b.FMT_S16, tr, seg, ovl)
rows = plan.segment_count(n)
scores = torch.randn((rows, C), device=dev) * 2 - 6
d_idx = torch.empty((rows, 5), dtype=torch.int32, device=dev)
d_conf = torch.empty((rows, 5), dtype=torch.float32, device=dev)
d_cnt = torch.empty((rows,), dtype=torch.int32, device=dev)
cfg = b.PostConfig(activation=b.ACT_SOFTMAX if config == "c3" else b.ACT_SIGMOID)
for _ in range(iters):
    r = plan.run(pcm, pad_to_batch=64)
    ctx.post_run_device(scores.data_ptr(), rows, C, rows, cfg, None, None, d_idx.data_ptr(), d_conf.data_ptr(), d_cnt.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    r = plan.run(pcm, pad_to_batch=64)
e1.record(); torch.cuda.synchronize()
print(f"{config}: {minutes} min audio, {r.nseg} windows: front end {e0.elapsed_time(e1)/iters:.3f} ms/iter "
      f"-> {minutes/60/(e0.elapsed_time(e1)/iters/1e3):.1f} audio-h/s")
e0.record()
for _ in range(iters):
    ctx.post_run_device(scores.data_ptr(), rows, C, rows, cfg, None, None, d_idx.data_ptr(), d_conf.data_ptr(), d_cnt.data_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{config}: post {rows} x {C}: {ms*1e3:.1f} us/iter -> {rows*C*4/ms/1e6:.0f} GB/s")
