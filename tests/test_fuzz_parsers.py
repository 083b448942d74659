"""Malformed WAV / FLAC files must come back as an error code (or be read correctly), never crash the host: the C ABI
promises error codes (include/birda_b200.h).  A bounded, seeded mutation run through ctypes; the same mutations run under
AddressSanitizer + UBSan with `tools/fuzz/run.sh` (20 000 iterations each, clean)."""
import ctypes as C
import os
import struct

import numpy as np

from birda_b200 import _lib
from birda_b200.synth import synth_pcm, write_wav
from tests import flac_enc

lib = _lib.lib


def test_wav_header_mutations(tmp_path):
    rng = np.random.default_rng(1)
    base = str(tmp_path / "a.wav")
    write_wav(base, synth_pcm(1, 0.05, 16_000, 2), 16_000, 2)
    raw = bytearray(open(base, "rb").read())
    path = str(tmp_path / "f.wav")
    accepted = rejected = 0
    for it in range(400):
        b = bytearray(raw)
        mode = it % 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, 44))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[: int(rng.integers(0, len(b)))]
        elif mode == 2:
            pos = int(rng.integers(0, 40))
            b[pos:pos + 4] = struct.pack("<I", int(rng.choice([0, 1, 2, 0x7FFFFFFF, 0xFFFFFFFF, 0xFFFFFFFE, len(b), len(b) + 1])))
        else:
            pos = int(rng.integers(12, 44))
            b[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8))
        open(path, "wb").write(b)
        info = _lib.WavInfo()
        rc = lib.bb_wav_probe(path.encode(), C.byref(info))
        if rc != 0:
            rejected += 1
            assert rc < 0
            continue
        accepted += 1
        # what the header promises must be inside the file
        bytes_per_frame = info.channels * (info.bits_per_sample // 8)
        assert bytes_per_frame > 0 and info.data_offset + info.frames * bytes_per_frame <= len(b)
        frames = min(int(info.frames), 500)
        dst = (C.c_ubyte * max(1, frames * bytes_per_frame))()
        assert lib.bb_wav_read(path.encode(), C.byref(info), 0, frames, dst) == 0
    assert accepted > 50 and rejected > 50


def test_flac_stream_mutations():
    rng = np.random.default_rng(2)
    pcm = synth_pcm(3, 0.5, 16_000, 2).reshape(-1, 2)
    data = bytearray(flac_enc.encode(pcm, 16_000, 16, style=dict(kinds=["lpc"], stereo="mid_side", part_order=2)))
    indexed = rejected = 0
    for it in range(400):
        b = bytearray(data)
        mode = it % 3
        if mode == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[: int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, 60))] = int(rng.integers(0, 256))
        arr = (C.c_ubyte * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")
        info = _lib.FlacInfo()
        if lib.bb_flac_probe_bytes(arr, len(b), C.byref(info)) != 0:
            rejected += 1
            continue
        cap = 256
        off, first, bs, n = (C.c_uint64 * cap)(), (C.c_uint64 * cap)(), (C.c_uint32 * cap)(), C.c_uint64()
        rc = lib.bb_flac_index(arr, len(b), C.byref(info), off, first, bs, cap, C.byref(n))
        if rc != 0:
            rejected += 1
            continue
        indexed += 1
        assert n.value <= cap and all(off[i] < len(b) for i in range(n.value))
        assert all(off[i] < off[i + 1] for i in range(n.value - 1))
    assert indexed > 50 and rejected > 20
