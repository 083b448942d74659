"""WAV / RF64 ingest (host code in the library, CPU only) + the file -> front end path on the GPU."""
import struct
import wave

import numpy as np
import pytest

import birda_b200 as b
from birda_b200.synth import synth_pcm


def write_wav(path, pcm, rate, channels, tag=1, bits=16, extensible=False, rf64=False, junk=True):
    raw = pcm.tobytes()
    fmt = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    if extensible:
        guid_tail = bytes.fromhex("000000001000800000aa00389b71")
        fmt += struct.pack("<HHI", 22, bits, 3) + struct.pack("<H", tag) + guid_tail
    chunks = b""
    if rf64:
        chunks += b"ds64" + struct.pack("<IQQQI", 28, 0, len(raw), len(raw) // (channels * bits // 8), 0)
    if junk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"hello" + b"\0"        # odd-sized chunk + pad byte
    chunks += b"fmt " + struct.pack("<I", len(fmt)) + fmt
    chunks += b"data" + struct.pack("<I", 0xFFFFFFFF if rf64 else len(raw)) + raw
    head = (b"RF64" + struct.pack("<I", 0xFFFFFFFF) if rf64 else b"RIFF" + struct.pack("<I", 4 + len(chunks))) + b"WAVE"
    open(path, "wb").write(head + chunks)


@pytest.mark.parametrize("dtype,tag,bits,fmt", [(np.int16, 1, 16, b.FMT_S16), (np.int32, 1, 32, b.FMT_S32), (np.float32, 3, 32, b.FMT_F32)])
@pytest.mark.parametrize("extensible,rf64", [(False, False), (True, False), (False, True)])
def test_probe_and_read(tmp_path, dtype, tag, bits, fmt, extensible, rf64):
    pcm = synth_pcm(5, 1.3, 22_050, 2, dtype)
    p = str(tmp_path / "a.wav")
    write_wav(p, pcm, 22_050, 2, tag, bits, extensible, rf64)
    info = b.wav_probe(p)
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.fmt) == (22_050, 2, bits, fmt)
    assert info.frames == pcm.size // 2
    assert np.array_equal(b.wav_read(p, info).view(np.uint8), pcm.view(np.uint8))
    part = b.wav_read(p, info, first_frame=1000, frames=777)
    assert np.array_equal(part.view(np.uint8), pcm[2000: 2000 + 2 * 777].view(np.uint8))


def test_python_wave_module_file(tmp_path):
    pcm = synth_pcm(6, 0.5, 48_000, 1)
    p = str(tmp_path / "w.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(48_000); w.writeframes(pcm.tobytes())
    info = b.wav_probe(p)
    assert (info.sample_rate, info.channels, info.fmt, info.frames) == (48_000, 1, b.FMT_S16, pcm.size)
    assert np.array_equal(b.wav_read(p, info), pcm)


def test_unsupported_and_errors(tmp_path):
    p = str(tmp_path / "u8.wav")
    write_wav(p, np.zeros(100, np.uint8), 8000, 1, 1, 8)
    with pytest.raises(b.BirdaError) as e:
        b.wav_probe(p)
    assert e.value.code == -4            # the reference drops U8 buffers silently (decode.rs:407-409)
    with pytest.raises(b.BirdaError) as e:
        b.wav_probe(str(tmp_path / "missing.wav"))
    assert e.value.code == -10
    q = str(tmp_path / "not.wav"); open(q, "wb").write(b"OggS" + b"\0" * 64)
    with pytest.raises(b.BirdaError):
        b.wav_probe(q)
    # truncated data chunk: decode what exists
    pcm = synth_pcm(7, 0.2, 16_000, 1)
    t = str(tmp_path / "t.wav"); write_wav(t, pcm, 16_000, 1)
    data = open(t, "rb").read(); open(t, "wb").write(data[:-101])
    info = b.wav_probe(t)
    assert info.frames == (pcm.size * 2 - 101) // 2


@pytest.mark.gpu
def test_file_to_segments_matches_oracle(tmp_path):
    from oracle import frontend as ofe
    pcm = synth_pcm(8, 7.3, 44_100, 2)
    p = str(tmp_path / "f.wav"); write_wav(p, pcm, 44_100, 2)
    info = b.wav_probe(p)
    ctx = b.Context(0)
    plan = b.FrontEndPlan(ctx, info.sample_rate, info.channels, info.fmt, 48_000, 144_000, 0)
    res = plan.run(b.wav_read(p, info)); ctx.sync()
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 0, precision="f64")
    got = res.torch().cpu().numpy()[: res.nseg]
    rms = np.sqrt(np.mean(ref.segments.astype(np.float64) ** 2, axis=1, keepdims=True))
    assert (np.abs(got - ref.segments) / np.maximum(np.abs(ref.segments), np.maximum(rms, 1e-30))).max() <= 1e-5
    assert np.array_equal(res.start_sample, ref.start_sample)
    ctx.close()
