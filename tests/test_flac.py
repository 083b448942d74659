"""FLAC ingest: the test-side encoder (tests/flac_enc.py) against the library's probe (CPU) and GPU decoder (bit-exact:
FLAC is lossless, the decoded PCM is the PCM that was encoded), then FLAC -> front end against the oracle on that PCM."""
import numpy as np
import pytest

import birda_b200 as b
from birda_b200.synth import synth_pcm
from tests import flac_enc


def test_probe_reads_streaminfo():
    pcm = synth_pcm(3, 0.5, 22_050, 2).reshape(-1, 2)
    data = flac_enc.encode(pcm, 22_050, 16, blocksize=1152)
    info = b.flac_probe(data)
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.frames) == (22_050, 2, 16, pcm.shape[0])
    assert info.fmt == b.FMT_S16 and info.max_block == 1152 and data[info.first_frame_offset: info.first_frame_offset + 2] == b"\xff\xf8"
    with pytest.raises(b.BirdaError) as e:
        b.flac_probe(b"RIFF" + bytes(60))
    assert e.value.code == -4


def test_encoder_self_check():
    """The encoder's own pieces: CRCs against known answers, UTF-8 numbers, zigzag."""
    assert flac_enc.crc8(b"123456789") == 0xF4                 # CRC-8 (poly 0x07): check value of the catalogue
    assert flac_enc.crc16(b"123456789") == 0xFEE8              # CRC-16/UMTS (poly 0x8005, init 0, no reflection)
    assert flac_enc.utf8_number(0x7F) == b"\x7f" and flac_enc.utf8_number(0x80) == b"\xc2\x80"
    assert flac_enc.utf8_number(0x7FF) == b"\xdf\xbf" and flac_enc.utf8_number(0x800) == b"\xe0\xa0\x80"
    assert list(flac_enc.zigzag(np.array([0, -1, 1, -2, 2]))) == [0, 1, 2, 3, 4]


def flac_index(data):
    import ctypes as C

    from birda_b200 import _lib
    info = b.flac_probe(data)
    n = C.c_uint64()
    _lib.check(_lib.lib.bb_flac_index(data, len(data), C.byref(info), None, None, None, 0, C.byref(n)))
    off = np.zeros(n.value, np.uint64); first = np.zeros(n.value, np.uint64); bs = np.zeros(n.value, np.uint32)
    _lib.check(_lib.lib.bb_flac_index(data, len(data), C.byref(info), off.ctypes.data_as(_lib.u64p), first.ctypes.data_as(_lib.u64p),
                                      bs.ctypes.data_as(_lib.u32p), n.value, C.byref(n)))
    return off, first, bs


@pytest.mark.parametrize("variable", [False, True])
def test_frame_index_finds_every_frame(variable):
    """The host's frame index (header CRC-8 + number continuity, search from a guessed offset with a bounded look-ahead and a
    safe fallback) against the offsets the encoder wrote — fixed and wildly varying block sizes, with the sync pattern
    planted inside the audio (0xFFF8 as a sample value) so that false candidates exist."""
    pcm = synth_pcm(5, 6.0, 44_100, 2).reshape(-1, 2).astype(np.int64)
    pcm[::97, 0] = -8                                             # 0xFFF8: a verbatim subframe carries the sync code itself
    style = dict(kinds=["verbatim", "fixed2", "lpc"], stereo="independent")
    if variable:
        style["variable_blocks"] = [4096, 192, 4608, 256, 16, 1152, 777]
    data, where = flac_enc.encode(pcm, 44_100, 16, blocksize=1024, style=style, return_frames=True)
    off, first, bs = flac_index(data)
    assert list(off) == [w[0] for w in where]
    assert int(bs.sum()) == pcm.shape[0] and first[0] == 0 and np.array_equal(first[1:], np.cumsum(bs)[:-1].astype(np.uint64))
    with pytest.raises(b.BirdaError):                             # a stream cut inside the metadata
        b.flac_probe(data[:30])


STYLES = [
    dict(kinds=["fixed0", "fixed1", "fixed2", "fixed3", "fixed4"], part_order=3),
    dict(kinds=["lpc"], lpc_order=8, lpc_precision=12, part_order=4, stereo="mid_side"),
    dict(kinds=["lpc", "fixed2"], lpc_order=32, lpc_precision=15, part_order=2, stereo="left_side", five_bit=True),
    dict(kinds=["verbatim", "fixed1"], stereo="right_side", escape_first=True, part_order=1),
    dict(kinds=["fixed2"], part_order=0, rate_in_header=False, bps_in_header=False),
    dict(kinds=["lpc"], lpc_order=1, lpc_precision=5, part_order=5, rate_explicit=True),
]


@pytest.mark.gpu
@pytest.mark.parametrize("style", range(len(STYLES)))
@pytest.mark.parametrize("rate,channels,bps,blocksize", [(44_100, 2, 16, 4096), (48_000, 1, 16, 1152), (96_000, 2, 24, 4608), (22_050, 2, 12, 300)])
def test_gpu_decode_is_bit_exact(style, rate, channels, bps, blocksize):
    rng = np.random.default_rng(100 * style + bps)
    n = int(rate * 0.9) + 17                                       # last block short
    base = synth_pcm(7 + style, n / rate + 0.01, rate, channels).reshape(-1, channels)[:n].astype(np.int64)
    if bps == 24:
        base = base * 256 + rng.integers(-128, 128, base.shape)     # real 24-bit content
    elif bps == 12:
        base = base >> 4
    base[1000:1600] = base[1000]                                    # a constant stretch
    base[5000:9000] &= ~0xF                                         # a stretch with wasted bits
    st = dict(STYLES[style])
    if style == 0:
        st["kinds"] = st["kinds"] + ["constant"]
    data = flac_enc.encode(base, rate, bps, blocksize=blocksize, style=st)
    ctx = b.Context(0)
    dec = b.FlacDecoder(ctx)
    out, info = dec.decode_to_numpy(data)
    assert (info.sample_rate, info.channels, info.bits_per_sample) == (rate, channels, bps)
    if bps <= 16:
        got = out.astype(np.int64).reshape(-1, channels) >> (16 - bps)
    else:
        u = out.reshape(-1, 3).astype(np.int64)
        v = u[:, 0] | (u[:, 1] << 8) | (u[:, 2] << 16)
        got = np.where(v >= 1 << 23, v - (1 << 24), v).reshape(-1, channels)
    assert got.shape == base.shape and np.array_equal(got, base)
    dec.close(); ctx.close()


@pytest.mark.gpu
def test_variable_blocksize_stream_and_long_file():
    """Blocking strategy 1 (the header carries the first sample's number, up to 36 bits in 7 bytes) with block sizes
    that change from frame to frame; long enough that sample numbers need 3- and 4-byte codes."""
    pcm = synth_pcm(17, 30.0, 44_100, 2).reshape(-1, 2).astype(np.int64)
    data = flac_enc.encode(pcm, 44_100, 16, style=dict(kinds=["fixed2", "lpc"], stereo="left_side", variable_blocks=[4096, 1152, 256, 4608, 777]))
    ctx = b.Context(0)
    dec = b.FlacDecoder(ctx)
    out, info = dec.decode_to_numpy(data)
    assert (info.min_block, info.max_block) == (256, 4608)
    assert np.array_equal(out.reshape(-1, 2).astype(np.int64), pcm)
    dec.close(); ctx.close()


@pytest.mark.gpu
def test_corrupt_frame_is_reported():
    pcm = synth_pcm(9, 1.0, 44_100, 2).reshape(-1, 2)
    data = bytearray(flac_enc.encode(pcm, 44_100, 16))
    data[len(data) // 2] ^= 0x10                                    # flip a bit inside some frame
    ctx = b.Context(0)
    dec = b.FlacDecoder(ctx)
    with pytest.raises(b.BirdaError) as e:
        dec.decode(bytes(data))
    assert e.value.code == -10 and ("CRC" in e.value.message or "frame" in e.value.message)
    dec.close(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("rate,channels,bps", [(44_100, 2, 16), (48_000, 1, 24)])
def test_flac_through_the_front_end(rate, channels, bps):
    """FLAC -> GPU decode -> K2 / K1 from the device-resident PCM equals the oracle on the PCM that was encoded
    (16-bit through the S16 arm; 24-bit through the S32 `<< 8` arm the reference takes for left-justified samples)."""
    from birda_b200.synth import pack_s24
    from oracle import frontend as ofe
    pcm = synth_pcm(21, 9.5, rate, channels).reshape(-1, channels).astype(np.int64)
    if bps == 24:
        pcm = pcm * 256 + 77
    data = flac_enc.encode(pcm, rate, bps, style=dict(kinds=["lpc", "fixed2"], stereo="mid_side"))
    ctx = b.Context(0)
    dec = b.FlacDecoder(ctx)
    ptr, frames, info = dec.decode(data)
    assert frames == pcm.shape[0]
    plan = b.FrontEndPlan(ctx, rate, channels, info.fmt, 48_000, 144_000, 72_000)
    res = plan.run(ptr, frames=frames, is_device=True); ctx.sync()
    host = pcm.astype(np.int16).reshape(-1) if bps == 16 else pack_s24(pcm.astype(np.int32).reshape(-1))
    ref = ofe.decode_and_stream(host, channels, rate, 48_000, 144_000, 72_000, precision="f64" if rate != 48_000 else "f32")
    got = res.torch().cpu().numpy()[: res.nseg]
    assert res.nseg == ref.segments.shape[0] and np.array_equal(res.start_sample, ref.start_sample)
    if rate == 48_000:
        assert np.array_equal(got, ref.segments)
    else:
        rms = np.sqrt(np.mean(ref.segments.astype(np.float64) ** 2, axis=1, keepdims=True))
        assert (np.abs(got - ref.segments) / np.maximum(np.abs(ref.segments), np.maximum(rms, 1e-30))).max() <= 1e-5
    plan.close(); dec.close(); ctx.close()


@pytest.mark.gpu
def test_flac_file_through_the_native_pipeline_equals_wav(tmp_path):
    """bb_pipeline_process_wav takes FLAC files too (by content): the same detections as the WAV of the same PCM, alone
    and through the two-worker pool."""
    from birda_b200.pipeline import NativePipeline, NativePool, ProcessingConfig
    from birda_b200.synth import write_wav
    C = 6522
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=1.5, batch_size=8, min_confidence=0.1)
    pcm = synth_pcm(33, 40.0, 44_100, 2)
    wav = str(tmp_path / "a.wav"); write_wav(wav, pcm, 44_100, 2)
    fl = str(tmp_path / "a.flac")
    open(fl, "wb").write(flac_enc.encode(pcm.reshape(-1, 2), 44_100, 16, style=dict(kinds=["lpc", "fixed3"], stereo="mid_side", part_order=4)))
    ctx = b.Context(0)
    nat = NativePipeline(ctx, cfg, b.StandIn(0, 144_000, C, 8, seed=3, stream=ctx.stream))
    key = lambda r: [(d.segment, d.index, round(d.confidence, 6), d.start_time, d.end_time) for d in r.detections]
    rw, rf = nat.process_wav(wav), nat.process_wav(fl)
    assert rw.segments == rf.segments and rw.effective_batch_size == rf.effective_batch_size
    assert key(rw) == key(rf) and len(rw.detections) > 10
    nat.close(); ctx.close()
    pool = NativePool([0, 0], [cfg, cfg], [b.StandIn(0, 144_000, C, 8, seed=3) for _ in range(2)])
    got = pool.process_wavs([fl, wav, fl])
    pool.close()
    assert key(got[0]) == key(got[1]) == key(got[2]) == key(rw)
