"""Build-time self check of the K2 fast-path templates: the same code the device runs, executed
lane by lane on the CPU against a naive double-precision DFT pipeline (CPU only, needs nvcc)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_k2_fast_templates_on_host(tmp_path):
    exe = str(tmp_path / "host_k2_check")
    r = subprocess.run(["nvcc", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                        "-diag-suppress", "20011,20014", "-o", exe, os.path.join(ROOT, "tests", "host_k2_check.cu")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "all plans ok" in r.stdout, r.stdout
