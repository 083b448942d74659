#!/usr/bin/env python3
"""Parity spot check for an A/B kernel variant (BIRDA_B200_LIB): a few C2 / C3 windows against the oracle.
usage: python tools/ab_check.py [c2|c3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import birda_b200 as b
from birda_b200.synth import synth_pcm
from oracle import frontend as ofe

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
sr, ch, tr, seg, ovl = (44_100, 2, 48_000, 144_000, 72_000) if cfg == "c2" else (48_000, 1, 32_000, 160_000, 0)
ctx = b.Context(0)
pcm = synth_pcm(2, 14.3, sr, ch)
plan = b.FrontEndPlan(ctx, sr, ch, b.FMT_S16, tr, seg, ovl)
res = plan.run(pcm, pad_to_batch=4)
ctx.sync()
got = res.torch().cpu().numpy()
ref = ofe.decode_and_stream(pcm, ch, sr, tr, seg, ovl, precision="f64")
rms = np.sqrt(np.mean(ref.segments.astype(np.float64) ** 2, axis=1, keepdims=True))
err = (np.abs(got[: res.nseg] - ref.segments) / np.maximum(np.abs(ref.segments), np.maximum(rms, 1e-30))).max()
print(f"{cfg}: {res.nseg} windows, parity err {err:.2e} {'OK' if err <= 1e-5 else 'FAIL'}")
