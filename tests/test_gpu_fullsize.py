"""Full BASELINE.json sizes on the GPU: EVERY row against the threaded C oracle (f32 restatement, all host cores),
spot rows against the f64 oracle, plus size-independent properties."""
import os

import numpy as np
import pytest

import birda_b200 as b
from birda_b200.pipeline import FilePipeline, ProcessingConfig
from birda_b200.shard import shard_files
from birda_b200.synth import synth_pcm
from oracle import frontend as ofe
from oracle import rules as orules

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    rms = np.sqrt(np.mean(ref.astype(np.float64) ** 2, axis=1, keepdims=True))
    return float((np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref), np.maximum(rms, 1e-30))).max())


def all_rows_err(out, ref, chunk=64):
    """Worst per-sample error over ALL rows (north_star gate: |got - ref| <= 1e-5 * max(|ref|, rms(ref row))),
    compared in row chunks so the f64 temporaries stay small.  `out` is the device tensor, `ref` the oracle's rows.
    Returns (worst error, its row)."""
    worst, worst_row = 0.0, -1
    for r0 in range(0, ref.shape[0], chunk):
        w = ref[r0: r0 + chunk].astype(np.float64)
        g = out[r0: r0 + w.shape[0]].cpu().numpy().astype(np.float64)
        rms = np.sqrt(np.mean(w ** 2, axis=1, keepdims=True))
        e = (np.abs(g - w) / np.maximum(np.abs(w), np.maximum(rms, 1e-30))).max(axis=1)
        i = int(e.argmax())
        if e[i] > worst:
            worst, worst_row = float(e[i]), r0 + i
    return worst, worst_row


def c_oracle_all_rows(pcm, ch, sr, tr, seg, ovl):
    from oracle import cport
    return cport.frontend(pcm, ch, sr, tr, seg, ovl, threads=os.cpu_count() or 1)


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0)
    yield c
    c.close()


def test_c2_full_hour_all_rows_and_tables(ctx):
    """C2: 1 h 44.1 kHz stereo, overlap 1.5 s, batch 64 -> 2400 windows (+32 padding rows).  All tables
    bit-exact; 24 rows spread over the hour (first, last, tail, pair boundaries) within 1e-5."""
    import torch
    sec = 3600
    base = synth_pcm(2, 60.0, 44_100, 2).reshape(-1, 2)
    pcm = np.tile(base, (sec // 60, 1))
    pcm[:: 9973, 0] += 17                                    # break the exact periodicity
    pcm = np.ascontiguousarray(pcm).reshape(-1)
    plan = b.FrontEndPlan(ctx, 44_100, 2, b.FMT_S16, 48_000, 144_000, 72_000)
    res = plan.run(pcm, pad_to_batch=64); ctx.sync()
    assert (res.nseg, res.rows) == (2400, 2432)
    rows = sorted(set([0, 1, 2, 3, 62, 63, 64, 65, 777, 1200, 1201, 2000, 2396, 2397, 2398, 2399]))
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 72_000, precision="f64", only=rows)
    assert np.array_equal(res.start_sample, ref.start_sample)
    assert res.start_time.tobytes() == ref.start_time.tobytes() and res.end_time.tobytes() == ref.end_time.tobytes()
    out = res.torch()
    got = out[torch.tensor(rows, device=out.device)].cpu().numpy()
    assert rel_err(got, ref.segments[rows].astype(np.float64)) <= 1e-5
    assert not bool(out[2400:].any())                        # batch padding is silence
    # every one of the 2400 rows against the threaded C oracle (f32 arithmetic like the reference)
    cseg, css, cst, cet = c_oracle_all_rows(pcm, 2, 44_100, 48_000, 144_000, 72_000)
    assert cseg.shape == (2400, 144_000) and np.array_equal(res.start_sample, css)
    assert res.start_time.tobytes() == cst.tobytes() and res.end_time.tobytes() == cet.tobytes()
    worst, row = all_rows_err(out, cseg)
    print(f"C2 all 2400 rows vs C oracle: worst {worst:.3e} at row {row}")
    assert worst <= 1e-5, (worst, row)
    del cseg
    tail = out[2399].cpu().numpy()
    # half-filled last window, zero padded: silence after the filter's tail.  The reference's transforms of all-zero
    # blocks give exact zeros from 72 000 + 2 240 on; K2's own blocking (4 480-sample transforms) leaves rounding noise
    # (~1e-9, inside the 1e-5 gate checked above) up to the end of the block that holds the last input sample
    assert tail[:70_000].any() and not tail[72_000 + 4_480 + 1_120:].any()
    assert float(np.abs(tail[72_000 + 2_240:]).max()) < 1e-7
    plan.close()


def test_c2_full_hour_device_resident_whole_and_split_items(ctx):
    """C2 from a device-resident buffer = ONE K2 launch over all 2432 rows: the first 1184 row pairs are whole-row work
    items, the last 32 pairs are cut into short runs with a recomputed carry block.  Rows of both kinds are checked."""
    import torch
    base = synth_pcm(5, 60.0, 44_100, 2).reshape(-1, 2)
    pcm = np.tile(base, (60, 1)); pcm[:: 9967, 1] -= 23
    pcm = np.ascontiguousarray(pcm).reshape(-1)
    d = torch.from_numpy(pcm).cuda()
    plan = b.FrontEndPlan(ctx, 44_100, 2, b.FMT_S16, 48_000, 144_000, 72_000)
    res = plan.run(d, pad_to_batch=64); ctx.sync()
    assert (res.nseg, res.rows) == (2400, 2432)
    rows = [0, 1, 500, 1183, 1184, 2366, 2367, 2368, 2369, 2383, 2390, 2398, 2399]
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 72_000, precision="f64", only=rows)
    assert np.array_equal(res.start_sample, ref.start_sample)
    out = res.torch()
    got = out[torch.tensor(rows, device=out.device)].cpu().numpy()
    assert rel_err(got, ref.segments[rows].astype(np.float64)) <= 1e-5
    assert not bool(out[2400:].any())
    cseg, css, _, _ = c_oracle_all_rows(pcm, 2, 44_100, 48_000, 144_000, 72_000)
    worst, row = all_rows_err(out, cseg)
    print(f"C2 (device-resident, one launch) all 2400 rows vs C oracle: worst {worst:.3e} at row {row}")
    assert np.array_equal(res.start_sample, css) and worst <= 1e-5, (worst, row)
    plan.close()


def test_forced_short_runs_match_whole_rows(ctx, monkeypatch):
    """Every row cut into runs of 5 blocks (carry recomputed 25 times per row) gives the same samples as whole rows
    to within the resampler tolerance."""
    pcm = synth_pcm(6, 20.0, 44_100, 2)
    import torch
    d = torch.from_numpy(pcm).cuda()
    plan = b.FrontEndPlan(ctx, 44_100, 2, b.FMT_S16, 48_000, 144_000, 72_000)
    monkeypatch.setenv("BIRDA_K2_ITEM_BLOCKS", "129")
    r = plan.run(d); ctx.sync()
    whole = r.torch().cpu().numpy().copy()
    monkeypatch.setenv("BIRDA_K2_ITEM_BLOCKS", "5")
    r = plan.run(d); ctx.sync()
    cut = r.torch().cpu().numpy().copy()
    monkeypatch.delenv("BIRDA_K2_ITEM_BLOCKS")
    assert whole.shape == cut.shape and whole.shape[0] >= 12
    assert rel_err(cut, whole.astype(np.float64)) <= 2e-6
    ref = ofe.decode_and_stream(pcm, 2, 44_100, 48_000, 144_000, 72_000, precision="f64")
    assert rel_err(cut[: ref.segments.shape[0]], ref.segments.astype(np.float64)) <= 1e-5
    plan.close()


def test_c3_full_hour_perch_all_rows(ctx):
    """C3: 1 h 48 kHz mono -> 32 kHz, 5 s windows, batch 128 -> 720 windows."""
    import torch
    base = synth_pcm(3, 60.0, 48_000, 1)
    pcm = np.tile(base, 60); pcm[:: 7919] += 11
    plan = b.FrontEndPlan(ctx, 48_000, 1, b.FMT_S16, 32_000, 160_000, 0)
    res = plan.run(pcm, pad_to_batch=128); ctx.sync()
    assert (res.nseg, res.rows) == (720, 768)
    rows = [0, 1, 127, 128, 360, 719]
    ref = ofe.decode_and_stream(pcm, 1, 48_000, 32_000, 160_000, 0, precision="f64", only=rows)
    assert np.array_equal(res.start_sample, ref.start_sample) and res.end_time.tobytes() == ref.end_time.tobytes()
    out = res.torch()
    assert rel_err(out[torch.tensor(rows, device=out.device)].cpu().numpy(), ref.segments[rows].astype(np.float64)) <= 1e-5
    cseg, css, cst, cet = c_oracle_all_rows(pcm, 1, 48_000, 32_000, 160_000, 0)
    assert cseg.shape == (720, 160_000) and np.array_equal(res.start_sample, css)
    assert res.start_time.tobytes() == cst.tobytes() and res.end_time.tobytes() == cet.tobytes()
    worst, row = all_rows_err(out, cseg)
    print(f"C3 all 720 rows vs C oracle: worst {worst:.3e} at row {row}")
    assert worst <= 1e-5 and not bool(out[720:].any()), (worst, row)
    plan.close()


def test_c4_bat_hour_checksum_property(ctx):
    """C4: 1 h 256 kHz mono bat mode -> 8534 windows of 144000 (overlap 36000), no resampling.  Bit-exact
    property at full size: every window equals the converted PCM slice, checked through per-row sums of the
    exact integer samples (an f64 checksum of checksums) and 12 rows compared element-wise."""
    import torch
    n = 3600 * 256_000
    base = synth_pcm(4, 30.0, 256_000, 1, bat=True)
    pcm = np.tile(base, 120)[:n].copy(); pcm[:: 10007] -= 5
    seg, ovl = b.rules.segment_samples(0.5625, 0.0, 256_000, bat_mode=True)
    plan = b.FrontEndPlan(ctx, 256_000, 1, b.FMT_S16, 256_000, seg, ovl)
    res = plan.run(pcm); ctx.sync()
    assert res.nseg == 8534
    out = res.torch()
    st, tk = b.rules.segment_table(n, seg, ovl)
    assert np.array_equal(res.start_sample, st)
    csum = np.concatenate([[0], np.cumsum(pcm.astype(np.int64))])
    want = (csum[(st + tk).astype(np.int64)] - csum[st.astype(np.int64)]).astype(np.float64) / 32768.0
    got = out.double().sum(dim=1).cpu().numpy()
    assert np.array_equal(got, want)                          # sums of exactly representable values are exact in f64
    for r in (0, 1, 4000, 8532, 8533):
        ref = np.zeros(seg, np.float32); ref[: int(tk[r])] = pcm[int(st[r]): int(st[r] + tk[r])].astype(np.float32) / np.float32(32768.0)
        assert np.array_equal(out[r].cpu().numpy(), ref)
    plan.close()


def test_c5_mixed_rate_directory_sharded(ctx):
    """C5 (100 of the 1000 files, 30-39 s instead of 10 min each): files cycling over the six rates and mono / stereo,
    BirdNET windows at overlap 0; the shard plan over 8 ranks covers every file once; every file's tables are
    bit-exact and EVERY row of every file matches the threaded C oracle to 1e-5 (same-rate files: bit-exact)."""
    rates = [16_000, 22_050, 32_000, 44_100, 48_000, 96_000]
    files = [(1000 + i, rates[i % 6], 1 + (i % 2), 30.0 + 3.0 * (i % 4)) for i in range(100)]
    shards = shard_files([f[3] for f in files], 8)
    assert sorted(i for part in shards for i in part) == list(range(100))
    plans, worst_all = {}, (0.0, None)
    for seed, sr, ch, dur in files:
        pcm = synth_pcm(seed, dur, sr, ch)
        if (sr, ch) not in plans:
            plans[(sr, ch)] = b.FrontEndPlan(ctx, sr, ch, b.FMT_S16, 48_000, 144_000, 0)
        res = plans[(sr, ch)].run(pcm, pad_to_batch=8); ctx.sync()
        cseg, css, cst, cet = c_oracle_all_rows(pcm, ch, sr, 48_000, 144_000, 0)
        assert res.nseg == cseg.shape[0] and np.array_equal(res.start_sample, css)
        assert res.start_time.tobytes() == cst.tobytes() and res.end_time.tobytes() == cet.tobytes()
        out = res.torch()
        if sr == 48_000:
            assert np.array_equal(out[: res.nseg].cpu().numpy(), cseg), (seed, sr, ch)
        else:
            worst, row = all_rows_err(out[: res.nseg], cseg)
            assert worst <= 1e-5, (seed, sr, ch, worst, row)
            if worst > worst_all[0]:
                worst_all = (worst, (seed, sr, ch, row))
        assert not bool(out[res.nseg:].any())
    print(f"C5 100 files, all rows vs C oracle: worst {worst_all[0]:.3e} at {worst_all[1]}")
    for p in plans.values():
        p.close()


def test_k2_linearity_and_shift_properties_bit_exact(ctx):
    """Size-independent properties of the resampling path, checked bit for bit on 10 minutes of C2 audio:
    (1) homogeneity — every operation between the PCM and the packed window is linear and scaling by two is exact in
    binary floating point, so K2(2 x) == 2 K2(x) exactly; (2) shift — window i of a file is window 0 of the file cut at
    i * hop: windows are independent of their neighbours and of which of the two packed streams carries them."""
    import torch
    pcm = (synth_pcm(8, 600.0, 44_100, 2).astype(np.int32) // 2).astype(np.int16)          # |x| <= 16384: 2x fits
    plan = b.FrontEndPlan(ctx, 44_100, 2, b.FMT_S16, 48_000, 144_000, 72_000)
    r1 = plan.run(pcm); ctx.sync()
    a = r1.torch()[: r1.nseg].clone()
    r2 = plan.run((pcm * 2).astype(np.int16)); ctx.sync()
    assert r2.nseg == r1.nseg == 400
    assert torch.equal(r2.torch()[: r2.nseg], a * 2.0)
    src_seg, hop = 132_300, 66_150
    for i in (1, 2, 7, 198, 333):
        cut = pcm[i * hop * 2: (i * hop + src_seg) * 2]
        r3 = plan.run(cut); ctx.sync()
        assert r3.nseg >= 1 and torch.equal(r3.torch()[0], a[i]), i
    plan.close()
