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


def test_24bit_pcm_probe_and_read(tmp_path):
    """24-bit PCM is the common field-recorder format: the reference converts it through its S32 arm
    (symphonia presents `sample << 8`; decode.rs:386-402), so the ingest reports FMT_S24 (3-byte packed)."""
    from birda_b200.synth import synth_pcm24
    pcm = synth_pcm24(9, 0.7, 48_000, 2)
    for extensible in (False, True):
        p = str(tmp_path / f"s24_{int(extensible)}.wav")
        write_wav(p, pcm, 48_000, 2, 1, 24, extensible)
        info = b.wav_probe(p)
        assert (info.sample_rate, info.channels, info.bits_per_sample, info.fmt) == (48_000, 2, 24, b.FMT_S24)
        assert info.frames == pcm.size // 6
        assert np.array_equal(b.wav_read(p, info), pcm)
        assert np.array_equal(b.wav_read(p, info, first_frame=101, frames=333), pcm[101 * 6: (101 + 333) * 6])


def test_inconsistent_headers_are_rejected(tmp_path):
    """block_align != channels * bytes, or EXTENSIBLE valid bits < container bits: reading at the nominal stride
    would decode garbage, so the file is refused (BB_ERR_UNSUPPORTED_FORMAT)."""
    pcm = synth_pcm(10, 0.1, 16_000, 2)
    p = str(tmp_path / "pad.wav"); write_wav(p, pcm, 16_000, 2, junk=False)
    data = bytearray(open(p, "rb").read())
    off = data.index(b"fmt ") + 8
    struct.pack_into("<H", data, off + 12, 6)                     # block_align 6 for 2 x 16-bit
    open(p, "wb").write(bytes(data))
    with pytest.raises(b.BirdaError) as e:
        b.wav_probe(p)
    assert e.value.code == -4 and "block size" in e.value.message
    q = str(tmp_path / "valid20.wav"); write_wav(q, synth_pcm(11, 0.1, 16_000, 1, np.int32), 16_000, 1, 1, 32, extensible=True, junk=False)
    data = bytearray(open(q, "rb").read())
    off = data.index(b"fmt ") + 8
    struct.pack_into("<H", data, off + 18, 20)                    # wValidBitsPerSample
    open(q, "wb").write(bytes(data))
    with pytest.raises(b.BirdaError) as e:
        b.wav_probe(q)
    assert e.value.code == -4


def test_streamed_writer_data_size_zero(tmp_path):
    """A data chunk size of 0 (a writer that never patched the header) means 'to the end of the file'."""
    pcm = synth_pcm(12, 0.25, 22_050, 1)
    p = str(tmp_path / "z.wav"); write_wav(p, pcm, 22_050, 1, junk=False)
    data = bytearray(open(p, "rb").read())
    struct.pack_into("<I", data, data.index(b"data") + 4, 0)
    open(p, "wb").write(bytes(data))
    info = b.wav_probe(p)
    assert info.frames == pcm.size
    assert np.array_equal(b.wav_read(p, info), pcm)


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 2, 3])
@pytest.mark.parametrize("tgt", [48_000, 32_000])
def test_24bit_file_through_the_front_end(tmp_path, channels, tgt):
    """24-bit WAV -> K1 (48 kHz, bit-exact) / K2 (48 -> 32 kHz, 1e-5) against the oracle's S32 `<< 8` conversion."""
    from birda_b200.synth import synth_pcm24
    from oracle import frontend as ofe
    pcm = synth_pcm24(13 + channels, 7.1, 48_000, channels)
    p = str(tmp_path / "f24.wav"); write_wav(p, pcm, 48_000, channels, 1, 24)
    info = b.wav_probe(p)
    seg = 144_000 if tgt == 48_000 else 160_000
    ctx = b.Context(0)
    plan = b.FrontEndPlan(ctx, info.sample_rate, info.channels, info.fmt, tgt, seg, seg // 2)
    res = plan.run(b.wav_read(p, info)); ctx.sync()
    ref = ofe.decode_and_stream(pcm, channels, 48_000, tgt, seg, seg // 2, precision="f64" if tgt != 48_000 else "f32")
    got = res.torch().cpu().numpy()[: res.nseg]
    assert res.nseg == ref.segments.shape[0] and np.array_equal(res.start_sample, ref.start_sample)
    if tgt == 48_000:
        assert np.array_equal(got, ref.segments)
    else:
        rms = np.sqrt(np.mean(ref.segments.astype(np.float64) ** 2, axis=1, keepdims=True))
        assert (np.abs(got - ref.segments) / np.maximum(np.abs(ref.segments), np.maximum(rms, 1e-30))).max() <= 1e-5
    ctx.close()


@pytest.mark.gpu
def test_24bit_odd_offsets_and_tail():
    """Device pointers at every byte alignment, windows that end inside the buffer's last word (byte-wise tail path)."""
    import torch
    from birda_b200.synth import synth_pcm24
    from oracle import frontend as ofe
    ctx = b.Context(0)
    for channels in (1, 2):
        pcm = synth_pcm24(21 + channels, 1.37, 48_000, channels)
        for off in (0, 1, 2, 3):
            buf = torch.zeros(pcm.size + 8, dtype=torch.uint8, device="cuda")
            buf[off: off + pcm.size] = torch.from_numpy(pcm).cuda()
            plan = b.FrontEndPlan(ctx, 48_000, channels, b.FMT_S24, 48_000, 14_404, 1_001)
            frames = pcm.size // (3 * channels)
            res = plan.run(buf.data_ptr() + off, frames=frames, is_device=True); ctx.sync()
            ref = ofe.decode_and_stream(pcm, channels, 48_000, 48_000, 14_404, 1_001)
            assert np.array_equal(res.torch().cpu().numpy()[: res.nseg], ref.segments), (channels, off)
            plan.close()
    ctx.close()


def test_parallel_read_equals_serial(tmp_path):
    """bb_wav_read_parallel: slices read by several threads land where one pread would put them (any thread count,
    ranges that do not divide evenly, reads below the slicing threshold)."""
    import ctypes as C

    from birda_b200 import _lib
    pcm = synth_pcm(31, 140.0, 48_000, 2)                          # 26.9 MB: more than a few 4 MB slices
    p = str(tmp_path / "big.wav"); write_wav(p, pcm, 48_000, 2)
    info = b.wav_probe(p)
    for first, frames in ((0, info.frames), (12_345, info.frames - 20_001), (7, 1000)):
        want = pcm[first * 2: (first + frames) * 2]
        for threads in (1, 2, 3, 5, 16):
            got = np.zeros(frames * 2, np.int16)
            rc = _lib.lib.bb_wav_read_parallel(p.encode(), C.byref(info._c) if hasattr(info, "_c") else C.byref(info), first, frames,
                                               got.ctypes.data_as(C.c_void_p), threads)
            assert rc == 0 and np.array_equal(got, want), (first, frames, threads)
