"""The piece table of a streamed file (csrc/pieces.hpp, used by csrc/pipeline.cpp) on the CPU: every piece but the last is
whole batches of full windows, pieces continue each other at the hop, the last one runs to the end of the file."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_piece_tables(tmp_path):
    exe = str(tmp_path / "host_pieces_check")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "host_pieces_check.cpp")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and " 0 bad" in r.stdout, r.stdout
