"""Per-file pipeline over the C ABI against the oracle pipeline with the same stand-in classifier."""
import numpy as np
import pytest

import birda_b200 as b
from birda_b200.pipeline import FilePipeline, ProcessingConfig
from birda_b200.synth import synth_pcm
from oracle import frontend as ofe
from oracle import post as opost
from oracle import rules as orules

pytestmark = pytest.mark.gpu


class StandInClassifier:
    """NOT BirdNET: a fixed random projection of 48 band energies, with the real I/O contract
    ([B, samples] f32 -> [B, C] f32 logits) so batching / padding / post can be exercised."""

    def __init__(self, samples: int, classes: int, seed: int = 0):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.w = (torch.randn(48, classes, generator=g) * 0.25).cuda()       # logits ~ N(-6, 2.5): a few classes clear 0.1
        self.bias = (torch.randn(classes, generator=g) * 1.0 - 6.0).cuda()
        self.frame = samples // 48

    def __call__(self, x):
        import torch
        e = x[:, : self.frame * 48].reshape(x.shape[0], 48, self.frame).double().pow(2).mean(dim=2)
        feat = torch.log10(e + 1e-6).float()
        return (feat @ self.w + self.bias).contiguous()


@pytest.mark.parametrize("sr,channels,overlap,batch,rerank", [(48_000, 1, 0.0, 8, False), (44_100, 2, 1.5, 16, True)])
def test_pipeline_matches_oracle(sr, channels, overlap, batch, rerank):
    import torch
    C = 6522
    ctx = b.Context(0)
    labels = [f"Genus{i} species{i}_Common {i}" for i in range(C)]
    rng = np.random.default_rng(3)
    mask = (rng.random(C) ** 2).astype(np.float32); mask[rng.choice(C, 305, replace=False)] = np.nan
    d_mask = torch.from_numpy(mask).cuda()
    clf = StandInClassifier(144_000, C)
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=overlap, batch_size=batch, min_confidence=0.1,
                           labels=labels, d_mask=d_mask.data_ptr(), rerank=rerank)
    pipe = FilePipeline(ctx, cfg, clf)
    pcm = synth_pcm(11, 61.0, sr, channels)
    res = pipe.process_pcm(pcm, channels, sr, b.FMT_S16)

    # oracle pipeline: oracle front end, same classifier on the oracle's segments, oracle post
    seg, ovl = orules.segment_and_overlap_samples(3.0, overlap, 48_000)
    ref = ofe.decode_and_stream(pcm, channels, sr, 48_000, seg, ovl)
    est = orules.estimate_segment_count(pcm.size // channels / sr, 3.0, overlap)
    B = orules.effective_batch_size(batch, est)
    assert res.effective_batch_size == B and res.segments == ref.segments.shape[0]
    assert res.batches == len(orules.batch_layout(res.segments, B))
    dets = []
    for first, valid, padded in orules.batch_layout(res.segments, B):
        xb = np.zeros((padded, seg), np.float32); xb[:valid] = ref.segments[first:first + valid]
        scores = clf(torch.from_numpy(xb).cuda()).cpu().numpy()
        rows = opost.post_process(scores, valid, opost.ACT_SIGMOID, 0.1, 5, mask, opost.FilterSettings(0.01, True, rerank))
        dets += opost.extract_detections(rows, ref.start_time, ref.end_time, labels, 0.1, first)
    dets = opost.sort_detections(dets)
    got = [(d.segment, d.index) for d in res.detections]
    want = [(d.segment, d.index) for d in dets]
    # the stand-in's scores depend on resampled samples (1e-5): allow boundary flips only
    assert len(set(got) ^ set(want)) <= max(3, len(want) // 20), (len(got), len(want))
    common = set(got) & set(want)
    gm = {(d.segment, d.index): d for d in res.detections}
    for d in dets:
        if (d.segment, d.index) in common:
            g = gm[(d.segment, d.index)]
            assert abs(g.confidence - float(d.confidence)) <= 1e-3
            assert np.float32(g.start_time) == d.start_time and np.float32(g.end_time) == d.end_time
            assert g.scientific_name == d.scientific_name
    st = [d.start_time for d in res.detections]
    assert st == sorted(st) and len(want) > 10
    pipe.close(); ctx.close()


def test_pipeline_bat_mode_no_resample():
    """Bat mode: 256 kHz fed raw, 144000-sample windows with 36000 overlap whatever --overlap says
    (src/pipeline/processor.rs:461-475, :502-508); the estimate still uses the user's overlap."""
    import torch
    ctx = b.Context(0)
    clf = StandInClassifier(144_000, 300, seed=3)
    cfg = ProcessingConfig(bat_mode=True, overlap=0.0, batch_size=32, min_confidence=0.05)
    pipe = FilePipeline(ctx, cfg, clf)
    pcm = synth_pcm(12, 4.1, 256_000, 1, bat=True)
    res = pipe.process_pcm(pcm, 1, 256_000, b.FMT_S16)
    ref = ofe.decode_and_stream(pcm, 1, 256_000, 256_000, 144_000, 36_000)
    assert res.segments == ref.segments.shape[0]
    est = orules.estimate_segment_count(pcm.size / 256_000, float(orules.BAT_SEGMENT_DURATION), 0.0)
    assert res.effective_batch_size == orules.effective_batch_size(32, est)
    ends = sorted({round(d.end_time - d.start_time, 6) for d in res.detections})
    assert ends == [0.5625] or not res.detections
    pipe.close(); ctx.close()


def test_native_cpp_pipeline_equals_python_mirror(tmp_path):
    """csrc/pipeline.cpp (bb_pipeline_*) against the Python mirror on PCM and on a WAV file streamed in pieces."""
    import torch
    from birda_b200.pipeline import NativePipeline
    from tests.test_wav_ingest import write_wav
    C = 6522
    ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(4)
    mask = (rng.random(C) ** 2).astype(np.float32); mask[rng.choice(C, 305, replace=False)] = np.nan
    d_mask = torch.from_numpy(mask).cuda()
    clf = StandInClassifier(144_000, C)
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=1.5, batch_size=8, min_confidence=0.1,
                           d_mask=d_mask.data_ptr())
    pcm = synth_pcm(13, 47.0, 44_100, 2)
    py = FilePipeline(ctx, cfg, clf).process_pcm(pcm, 2, 44_100, b.FMT_S16)
    nat = NativePipeline(ctx, cfg, clf)
    r1 = nat.process_pcm(pcm, 2, 44_100, b.FMT_S16)
    key = lambda r: [(d.segment, d.index, round(d.confidence, 6), d.start_time, d.end_time) for d in r.detections]
    assert r1.segments == py.segments and r1.effective_batch_size == py.effective_batch_size
    assert key(r1) == key(py) and len(r1.detections) > 10
    # the same audio from a WAV file, streamed in small pieces (several pieces, batches stay aligned)
    p = str(tmp_path / "f.wav"); write_wav(p, pcm, 44_100, 2)
    r2 = nat.process_wav(p, piece_frames=44_100 * 9)
    assert r2.segments == py.segments and key(r2) == key(py)
    nat.close(); ctx.close()


def test_native_pipeline_mixed_rate_directory_reuses_plans():
    """BASELINE config 5 in miniature: files cycling over rates and channel counts through ONE native pipeline.
    Every kind of file builds its plan once; results equal a fresh pipeline per file."""
    import time

    import torch
    from birda_b200.pipeline import NativePipeline
    C = 265
    ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    clf = StandInClassifier(144_000, C)
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=0.0, batch_size=4, min_confidence=0.1)
    kinds = [(16_000, 1), (22_050, 2), (32_000, 1), (44_100, 2), (48_000, 1), (96_000, 2)]
    files = [(sr, ch, synth_pcm(200 + i, 10.0, sr, ch)) for i, (sr, ch) in enumerate(kinds * 3)]
    key = lambda r: [(d.segment, d.index, round(d.confidence, 6), d.start_time, d.end_time) for d in r.detections]
    nat = NativePipeline(ctx, cfg, clf)
    t_first, t_again, results = 0.0, 0.0, []
    for i, (sr, ch, pcm) in enumerate(files):
        t0 = time.perf_counter()
        results.append(nat.process_pcm(pcm, ch, sr, b.FMT_S16))
        dt = time.perf_counter() - t0
        if i < len(kinds):
            t_first += dt
        else:
            t_again += dt / 2
    assert nat.plans_created == len(kinds)                     # 18 files, 6 kinds, 6 plans
    assert t_again < t_first                                   # later passes do not pay for plan creation
    for i in (0, 3, 7, 11, 17):                                # same detections as a pipeline that sees only this file
        sr, ch, pcm = files[i]
        one = NativePipeline(ctx, cfg, clf)
        r = one.process_pcm(pcm, ch, sr, b.FMT_S16)
        assert r.segments == results[i].segments == 4 and key(r) == key(results[i])
        one.close()
    nat.close(); ctx.close()


def test_native_pool_two_workers_match_single_pipeline(tmp_path):
    """csrc/pool.cpp: two worker threads (two contexts on GPU 0 when the box has one GPU, GPUs 0 and 1 otherwise) share a
    queue of WAV files, longest first; every file's detections equal those of a single native pipeline."""
    import torch
    from birda_b200.api import device_count
    from birda_b200.pipeline import NativePipeline, NativePool
    from tests.test_wav_ingest import write_wav
    C = 265
    devices = [0, 1] if device_count() > 1 else [0, 0]
    clfs = []
    for d in devices:
        with torch.cuda.device(d):
            clfs.append(StandInClassifier(144_000, C))
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=1.5, batch_size=4, min_confidence=0.1)
    paths, pcms = [], []
    for i, (sr, ch, sec) in enumerate([(44_100, 2, 21.0), (48_000, 1, 7.0), (32_000, 1, 33.0), (44_100, 2, 12.5), (96_000, 2, 9.0), (22_050, 1, 15.0)]):
        pcm = synth_pcm(300 + i, sec, sr, ch)
        p = str(tmp_path / f"f{i}.wav"); write_wav(p, pcm, sr, ch)
        paths.append(p); pcms.append((pcm, ch, sr))
    pool = NativePool(devices, [cfg] * len(devices), clfs)
    got = pool.process_wavs(paths)
    pool.close()
    ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    one = NativePipeline(ctx, cfg, clfs[0])
    key = lambda r: [(d.segment, d.index, round(d.confidence, 5), d.start_time, d.end_time) for d in r.detections]
    for p, (pcm, ch, sr), g in zip(paths, pcms, got):
        r = one.process_pcm(pcm, ch, sr, b.FMT_S16)
        assert g.segments == r.segments and g.effective_batch_size == r.effective_batch_size
        assert key(g) == key(r), p
    assert sum(len(g.detections) for g in got) > 20
    one.close(); ctx.close()


def test_native_pipeline_batch_hooks_bracket_every_batch():
    """The watchdog seam (src/pipeline/processor.rs:263-277): before/after fire once per batch, in order, with the
    padded batch size, the valid rows and the batch's first segment."""
    import torch
    from birda_b200.pipeline import NativePipeline
    ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    clf = StandInClassifier(144_000, 265)
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=0.0, batch_size=4, min_confidence=0.1)
    nat = NativePipeline(ctx, cfg, clf)
    events = []
    nat.set_batch_hooks(lambda rows, valid, first: events.append(("before", rows, valid, first)),
                        lambda rows, valid, first: events.append(("after", rows, valid, first)))
    res = nat.process_pcm(synth_pcm(40, 31.0, 48_000, 1), 1, 48_000, b.FMT_S16)        # 11 windows: 4 + 4 + 3 (padded)
    assert res.segments == 11
    want = []
    for first, valid in ((0, 4), (4, 4), (8, 3)):
        want += [("before", 4, valid, first), ("after", 4, valid, first)]
    assert events == want
    nat.close(); ctx.close()


def test_native_pipeline_batch_timeout_fails_the_file():
    """A classifier that outlives the library-kept watchdog: on_fire runs on the watchdog thread with the reference's
    arguments (whole seconds, batch size) and the file fails with BB_ERR_TIMEOUT; with a generous timeout the same
    pipeline then completes."""
    import time

    import torch
    from birda_b200.pipeline import NativePipeline
    ctx = b.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    inner = StandInClassifier(144_000, 265)
    slow = {"sleep": 0.6}

    def clf(x):
        time.sleep(slow["sleep"])
        return inner(x)

    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=0.0, batch_size=4, min_confidence=0.1)
    nat = NativePipeline(ctx, cfg, clf)
    fired = []
    nat.set_batch_timeout(150, lambda secs, batch: fired.append((secs, batch)))
    pcm = synth_pcm(41, 13.0, 48_000, 1)
    with pytest.raises(b.BirdaError) as e:
        nat.process_pcm(pcm, 1, 48_000, b.FMT_S16)
    assert e.value.code == -11 and "watchdog" in e.value.message
    assert fired == [(0, 4)]
    slow["sleep"] = 0.0
    nat.set_batch_timeout(60_000, lambda secs, batch: fired.append((secs, batch)))
    res = nat.process_pcm(pcm, 1, 48_000, b.FMT_S16)
    assert res.segments == 5 and len(fired) == 1
    nat.set_batch_timeout(0)
    assert nat.process_pcm(pcm, 1, 48_000, b.FMT_S16).segments == 5
    nat.close(); ctx.close()


def test_standin_classifier_matches_its_formula():
    """bb_standin (the library's stand-in classifier, NOT a model): logits = log10(mean square of 48 frames + 1e-6) @ W + b."""
    import torch
    st = b.StandIn(0, 144_000, 6522, 16, seed=5)
    W, bias = st.weights()
    x = torch.from_numpy(synth_pcm(50, 3.0 * 7, 48_000, 1).astype(np.float32).reshape(7, 144_000) / 32768.0).cuda().contiguous()
    got = st(x).cpu().numpy()
    xs = x.cpu().numpy().astype(np.float64).reshape(7, 48, 3000)
    feat = np.log10((xs ** 2).mean(axis=2) + 1e-6)
    want = feat @ W.astype(np.float64) + bias
    assert got.shape == (7, 6522) and np.abs(got - want).max() < 2e-4
    assert st.launches == 2
    frac = float((1 / (1 + np.exp(-want)) >= 0.1).mean())
    assert 1e-4 < frac < 0.05                                   # a few classes per window clear the usual threshold
    st.close()


def test_native_callback_multi_piece_file_and_pool(tmp_path):
    """The library's own bb_classify_fn (no Python in the loop): a WAV streamed in several pieces (reader thread + two
    pinned buffers, one result copy per piece) gives the same detections as the whole file in one piece and as the
    Python-callback pipeline; two pool workers on the same files agree too."""
    import torch
    from birda_b200.pipeline import NativePipeline, NativePool
    from tests.test_wav_ingest import write_wav
    C = 6522
    cfg = ProcessingConfig(target_rate=48_000, segment_duration=3.0, overlap=1.5, batch_size=8, min_confidence=0.1)
    pcm = synth_pcm(51, 95.0, 44_100, 2)
    p = str(tmp_path / "long.wav"); write_wav(p, pcm, 44_100, 2)
    ctx = b.Context(0)
    st = b.StandIn(0, 144_000, C, 8, seed=9, stream=ctx.stream)
    nat = NativePipeline(ctx, cfg, st)
    key = lambda r: [(d.segment, d.index, round(d.confidence, 6), d.start_time, d.end_time) for d in r.detections]
    whole = nat.process_wav(p)
    pieces = nat.process_wav(p, piece_frames=44_100 * 11)        # 9 pieces
    assert whole.segments == pieces.segments == 64 and key(whole) == key(pieces) and len(whole.detections) > 20
    frompcm = nat.process_pcm(pcm, 2, 44_100, b.FMT_S16)
    assert key(frompcm) == key(whole)
    # the same classifier behind the Python trampoline
    st2 = b.StandIn(0, 144_000, C, 8, seed=9)
    py = NativePipeline(ctx, cfg, lambda x: st2(x))
    assert key(py.process_wav(p, piece_frames=44_100 * 17)) == key(whole)
    # hooks force the per-batch synchronous path: same detections
    nat.set_batch_hooks(lambda *a: None, lambda *a: None)
    assert key(nat.process_wav(p, piece_frames=44_100 * 11)) == key(whole)
    nat.close(); py.close(); ctx.close()
    # pool: two workers, native callbacks
    paths = [p]
    for i, (sr, ch, sec) in enumerate([(48_000, 1, 20.0), (22_050, 1, 41.0), (32_000, 2, 15.0)]):
        q = str(tmp_path / f"g{i}.wav"); write_wav(q, synth_pcm(60 + i, sec, sr, ch), sr, ch); paths.append(q)
    sts = [b.StandIn(0, 144_000, C, 8, seed=9) for _ in range(2)]
    pool = NativePool([0, 0], [cfg, cfg], sts)
    got = pool.process_wavs(paths)
    pool.close()
    assert key(got[0]) == key(whole) and all(g.segments > 0 for g in got)
    ctx = b.Context(0)
    one = NativePipeline(ctx, cfg, b.StandIn(0, 144_000, C, 8, seed=9, stream=ctx.stream))
    for q, g in zip(paths[1:], got[1:]):
        assert key(one.process_wav(q)) == key(g), q
    one.close(); ctx.close()
