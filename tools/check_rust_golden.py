#!/usr/bin/env python3
"""Hold this repo's oracle (and, when a GPU is present, the CUDA path) to golden vectors dumped from the REFERENCE by
rust/examples/dump_resample_golden.rs on a machine with a Rust toolchain.  This is the step that would turn
"sample-level parity unpinned"  into "pinned"; it cannot run in the build image (no cargo, no rubato).

usage: python tools/check_rust_golden.py DIR          DIR holds cases.txt and <case>.in.f32 / <case>.out.f32"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import frontend as ofe


def lcg_signal(n: int, rate: int, seed: int) -> np.ndarray:
    """Bit-identical twin of lcg_signal in the Rust dumper (used to detect a corrupted or mismatched dump)."""
    out = np.empty(n, np.float32)
    state = seed
    mask = (1 << 64) - 1
    for i in range(n):
        state = (state * 6364136223846793005 + 1442695040888963407) & mask
        noise = ((state >> 40) / float(1 << 24) - 0.5) * 0.0632
        t = i / float(rate)
        s = 0.08 * np.sin(2.0 * np.pi * 1000.0 * t) + 0.08 * np.sin(2.0 * np.pi * 3217.0 * t) + 0.08 * np.sin(2.0 * np.pi * 7919.0 * t) + noise
        out[i] = np.float32(s)
    return out


def main(d: str) -> int:
    bad = 0
    for k, line in enumerate(open(os.path.join(d, "cases.txt"))):
        name, fr, to, n_in, n_out = line.split()
        fr, to, n_in, n_out = int(fr), int(to), int(n_in), int(n_out)
        x = np.fromfile(os.path.join(d, name + ".in.f32"), "<f4")
        y = np.fromfile(os.path.join(d, name + ".out.f32"), "<f4")
        assert x.size == n_in and y.size == n_out, name
        if n_in <= 20_000 and not np.array_equal(x, lcg_signal(n_in, fr, 1000 + k)):
            print(f"{name}: input differs from the generator (platform sin()?) - using the dumped input as is")
        got = ofe.resample(x, fr, to, precision="f64")
        if got.size != y.size:
            print(f"{name}: LENGTH {got.size} vs reference {y.size}"); bad += 1; continue
        rms = float(np.sqrt(np.mean(y.astype(np.float64) ** 2)))
        err = float((np.abs(got - y) / np.maximum(np.abs(y), max(rms, 1e-30))).max())
        ok = err <= 1e-5
        bad += 0 if ok else 1
        print(f"{name}: {n_in} -> {n_out} samples, oracle vs reference max rel err {err:.3e} {'OK' if ok else 'FAIL'}")
    print("parity with the reference's resampler:", "PINNED" if bad == 0 else f"{bad} case(s) differ - adjust oracle/frontend.py (window form, cutoff, tap normalisation: SURVEY.md 7.3 item 1)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
