#!/usr/bin/env python3
"""K6 timing: decode of a FLAC stream made by the test-side encoder (60 s of 44.1 kHz stereo, tiled to `minutes`)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import birda_b200 as b
from birda_b200.synth import synth_pcm
from tests import flac_enc

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
pcm = synth_pcm(3, 60.0, 44_100, 2).reshape(-1, 2)
pcm = np.tile(pcm, (int(minutes), 1))
t0 = time.perf_counter()
one = flac_enc.encode(pcm[: 44_100 * 60], 44_100, 16, style=dict(kinds=["lpc"], stereo="mid_side", part_order=4))
print(f"encoded 60 s in {time.perf_counter() - t0:.1f} s: {len(one)} bytes = {len(one) / (44_100 * 60 * 4):.2f} of the PCM")
data = flac_enc.encode(pcm, 44_100, 16, style=dict(kinds=["lpc"], stereo="mid_side", part_order=4)) if minutes <= 3 else None
if data is None:
    # frames are independent: repeat the minute's frames with renumbered headers is not valid FLAC, so encode once at length
    data = flac_enc.encode(pcm, 44_100, 16, style=dict(kinds=["fixed2"], stereo="mid_side", part_order=4))
ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
dec = b.FlacDecoder(ctx)
for _ in range(2):
    dec.decode(data)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    ptr, frames, info = dec.decode(data)
dt = (time.perf_counter() - t0) / n
print(f"decode {minutes} min stereo ({len(data) / 1e6:.1f} MB FLAC -> {frames * 4 / 1e6:.1f} MB PCM): {dt * 1e3:.2f} ms per file incl. host index + H2D "
      f"-> {minutes / 60 / dt:.0f} audio-h/s")

# the same file through the per-file pipeline (pinned staging, parallel read) next to its WAV
from birda_b200.pipeline import NativePipeline, ProcessingConfig
from birda_b200.synth import write_wav
os.makedirs("/dev/shm/bbflac", exist_ok=True)
open("/dev/shm/bbflac/a.flac", "wb").write(data); write_wav("/dev/shm/bbflac/a.wav", pcm.reshape(-1).astype(np.int16), 44_100, 2)
cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=0.0, batch_size=64, min_confidence=0.1)
nat = NativePipeline(ctx, cfg, b.StandIn(0, 144_000, 6522, 64, seed=1, stream=ctx.stream))
for name in ("a.wav", "a.flac"):
    nat.process_wav("/dev/shm/bbflac/" + name)
    t0 = time.perf_counter()
    for _ in range(5):
        r = nat.process_wav("/dev/shm/bbflac/" + name)
    dt = (time.perf_counter() - t0) / 5
    print(f"pipeline {name}: {dt * 1e3:.2f} ms per {minutes}-min file ({r.segments} windows) -> {minutes / 60 / dt:.0f} audio-h/s")
