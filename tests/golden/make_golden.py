#!/usr/bin/env python3
"""Generates the golden fixtures in this directory from the oracle (TEST INFRASTRUCTURE).

The reference cannot be built or imported in this environment (no Rust toolchain, un-vendored crates, no ONNX
Runtime: SURVEY.md 0 F10 / 8c), so these vectors are the ORACLE's outputs on seeded inputs, frozen so that a drift
of either the oracle or the CUDA path is caught; they are not outputs of the reference itself.  Once a Rust
toolchain is available the same inputs can be pushed through `birda::audio::resample` / `process_file` and the
`expected_*` arrays replaced.

usage: python tests/golden/make_golden.py   (from the repo root; rewrites tests/golden/*.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from birda_b200.synth import synth_logits, synth_pcm  # noqa: E402
from oracle import frontend as ofe  # noqa: E402
from oracle import melspec as om  # noqa: E402
from oracle import post as opost  # noqa: E402
from oracle import rules as orules  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def frontend_case(name, seed, seconds, sr, ch, tr, seg, ovl):
    pcm = synth_pcm(seed, seconds, sr, ch)
    r = ofe.decode_and_stream(pcm, ch, sr, tr, seg, ovl, precision="f64")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm=pcm, params=np.array([sr, ch, tr, seg, ovl], np.int64),
                        expected_segments=r.segments.astype(np.float32), expected_start_sample=r.start_sample,
                        expected_start_time=r.start_time, expected_end_time=r.end_time)


def main():
    # front end: short windows keep the files small; the block structure of the resampler is exercised in full
    frontend_case("frontend_pack_48k_mono", 101, 0.61, 48_000, 1, 48_000, 9_600, 2_400)
    frontend_case("frontend_pack_bat_256k", 102, 0.11, 256_000, 1, 256_000, 14_400, 3_600)
    frontend_case("frontend_resample_44k1_stereo_to_48k", 103, 0.55, 44_100, 2, 48_000, 9_600, 4_800)
    frontend_case("frontend_resample_48k_to_32k", 104, 0.70, 48_000, 1, 32_000, 8_000, 0)
    frontend_case("frontend_resample_96k_stereo_to_48k", 105, 0.45, 96_000, 2, 48_000, 9_600, 0)
    # post step
    x = synth_logits(106, 24, 265)
    rng = np.random.default_rng(106)
    mask = (rng.random(265) ** 2).astype(np.float32)
    mask[rng.choice(265, 20, replace=False)] = np.nan
    cases = {}
    for tag, kw in {"plain": {}, "mask_keep": dict(mask=mask, settings=opost.FilterSettings(0.01, True, False)),
                    "mask_rerank": dict(mask=mask, settings=opost.FilterSettings(0.01, False, True))}.items():
        rows = opost.post_process(x, 24, opost.ACT_SIGMOID, 0.1, 5, kw.get("mask"), kw.get("settings"))
        idx = np.full((24, 5), -1, np.int64); conf = np.zeros((24, 5), np.float32); cnt = np.zeros(24, np.int64)
        for r, row in enumerate(rows):
            cnt[r] = len(row)
            for j, (i, c) in enumerate(row):
                idx[r, j] = i; conf[r, j] = c
        cases[f"expected_{tag}_index"] = idx; cases[f"expected_{tag}_conf"] = conf; cases[f"expected_{tag}_count"] = cnt
    np.savez_compressed(os.path.join(HERE, "post_sigmoid_265.npz"), scores=x, mask=mask, **cases)
    # segment tables (start_sample, take) for awkward totals
    tabs = {}
    for total, seg, ovl in [(0, 100, 0), (1, 100, 0), (250, 100, 0), (250, 100, 30), (330, 100, 99), (1000, 300, 150), (144_001, 144_000, 72_000)]:
        t = orules.next_segment_table(total, seg, ovl)
        tabs[f"t_{total}_{seg}_{ovl}"] = np.array([(w.start_sample, w.take) for w in t], np.int64).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "segment_tables.npz"), **tabs)
    # spectrogram prefix
    seg = (synth_pcm(107, 0.5, 48_000, 1).astype(np.float32) / 32768.0).reshape(2, 12_000)
    w, mw = om.hann(512), om.mel_filterbank(32, 512, 48_000, 200.0, 12_000.0)
    np.savez_compressed(os.path.join(HERE, "melspec_512_32.npz"), segments=seg, window=w, mel_weights=mw,
                        params=np.array([512, 200, 58], np.int64), expected=om.melspec(seg, 512, 200, 58, w, mw).astype(np.float32))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
