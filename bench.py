#!/usr/bin/env python3
"""bench.py — audio-hours/sec of the birda front-end + post-inference hot path on B200.

A step = one pass of the hot path over one audio-hour of BASELINE config 2 per GPU:
  BirdNET v2.4, 1 h synthetic 44.1 kHz stereo s16 PCM -> downmix -> per-window 44.1->48 kHz FFT
  resample -> 2400 overlapped 3 s windows (overlap 1.5 s) packed as 38 batches of 64
  ([2432, 144000] f32), then sigmoid -> top-5 -> min_conf 0.1 -> range mask -> threshold over
  [2432, 6522] scores.  The model forward (ONNX Runtime, not part of this path and not
  available here) is not executed; scores are synthetic and resident in HBM.

`value`  : inputs resident in HBM (device timed, CUDA events, max over ranks).
`e2e`    : same metric through the C ABI with HOST (pinned) PCM: H2D copy + kernels + D2H of
           the detections inside the timed region.
`--impl reference`: the reference's CPU algorithm (oracle restatement) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SRC_RATE, TGT_RATE, CHANNELS = 44_100, 48_000, 2
SEG, OVL, BATCH = 144_000, 72_000, 64
SECONDS = 3600
CLASSES = 6522
WORKLOAD = ("C2: BirdNET v2.4 front end + post on 1 h synthetic 44.1 kHz stereo s16 per GPU per step "
            "(downmix + 44.1->48 kHz per-window FFT resample, overlap 1.5 s, batch 64, top-5, min_conf 0.1, range mask)")
ALGO_BYTES_FRONT = SECONDS * SRC_RATE * CHANNELS * 2 + 2400 * SEG * 4          # 2017.44 MB
ALGO_FLOPS_FRONT = 118_955 * 129 * 2400                                      # 36.8 GFLOP


def k2_profile(kernel: str):
    """Numbers that only a profiler can give (DRAM bytes per launch, shared-memory pipe share) for the dominant kernel,
    from the committed ncu --set full capture (profiles/k2_traffic.json, written by tools/collect_profiles.py).  They are
    quoted only when that capture names the kernel this run launched (bb_plan_describe); otherwise None + the reason."""
    p = os.path.join(ROOT, "profiles", "k2_traffic.json")
    try:
        j = json.load(open(p))
    except Exception:
        return None, None, "no committed capture"
    if j.get("kernel") != kernel:
        return None, None, f"committed capture is of another kernel ({j.get('kernel')})"
    return int(j["dram_bytes_per_launch"]), j.get("smem_pipe_pct"), j.get("source")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload_gpu(torch, device, seed):
    """1 h of 44.1 kHz stereo s16 made on the device (chirps + noise), plus scores and a range mask."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    n = SECONDS * SRC_RATE
    pcm = torch.empty((n, CHANNELS), dtype=torch.int16, device=device)
    blk = 60 * SRC_RATE
    f0 = torch.rand(5, generator=g, device=device) * 11_500 + 500
    f1 = torch.rand(5, generator=g, device=device) * 11_500 + 500
    for s in range(0, n, blk):
        e = min(n, s + blk)
        t = torch.arange(s, e, device=device, dtype=torch.float64) / SRC_RATE
        tt = torch.remainder(t, 10.0)
        x = torch.zeros(e - s, device=device, dtype=torch.float64)
        for i in range(5):
            x += 0.05 * torch.sin(2 * torch.pi * (f0[i].double() * tt + 0.5 * (f1[i] - f0[i]).double() / 10.0 * tt * tt))
        x += 0.0316 * torch.randn(e - s, generator=g, device=device, dtype=torch.float64)
        y = 0.8 * torch.roll(x, 7) + 0.0316 * torch.randn(e - s, generator=g, device=device, dtype=torch.float64)
        pcm[s:e, 0] = torch.clamp(torch.round(x * 32767), -32768, 32767).to(torch.int16)
        pcm[s:e, 1] = torch.clamp(torch.round(y * 32767), -32768, 32767).to(torch.int16)
    rows = 2432
    scores = torch.randn((rows, CLASSES), generator=g, device=device) * 2.0 - 6.0
    hot = torch.randint(0, CLASSES, (rows, 4), generator=g, device=device)
    scores.scatter_(1, hot, torch.randn((rows, 4), generator=g, device=device) + 2.0)
    mask = torch.rand(CLASSES, generator=g, device=device) ** 2
    mask[torch.randperm(CLASSES, generator=g, device=device)[:305]] = float("nan")
    return pcm.reshape(-1), scores, mask


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One bounded sample of the workload through the oracle (reference algorithm restated)."""
    seed, nseg = args
    from birda_b200.synth import synth_logits, synth_pcm
    from oracle import cport
    from oracle import frontend as ofe
    seconds = (nseg + 1) * 1.5
    pcm = synth_pcm(seed, seconds, SRC_RATE, CHANNELS)
    x = synth_logits(seed, nseg, CLASSES, adversarial=False)
    rng = np.random.default_rng(seed)
    mask = (rng.random(CLASSES) ** 2).astype(np.float32)
    mask[rng.choice(CLASSES, 305, replace=False)] = np.nan
    t0 = time.perf_counter()
    r = ofe.decode_and_stream(pcm, CHANNELS, SRC_RATE, TGT_RATE, SEG, OVL, batched=True)   # all blocks of a window in one pocketfft call
    n = r.segments.shape[0]
    cport.post(x[:n], min(n, nseg), 1, 0.1, 5, mask, None, 0.01, True, False, threads=1)
    dt = time.perf_counter() - t0
    return dt, seconds


def cpu_arm(cores: int, nseg_per_worker: int, seed: int = 100):
    """audio-hours/sec of the oracle over `cores` worker processes, each on its own sample."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker((seed, nseg_per_worker))]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(seed + i, nseg_per_worker) for i in range(cores)])
    wall = max(r[0] for r in res) if cores > 1 else res[0][0]
    audio_s = sum(r[1] for r in res)
    return audio_s / 3600.0 / wall, wall, time.perf_counter() - t0


CPU_SAMPLE_NOTE = ("oracle restatement of the reference's CPU path: pocketfft (SIMD) f32 block transforms, all blocks of a "
                   "window per call, + C post step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) or 1        # the threads this process may actually use
    nseg = 400                                   # 10 min of audio per worker per step
    vals = []
    for i in range(args.warmup + args.steps):
        v, wall, _ = cpu_arm(cores, nseg, seed=100 + 1000 * i)
        if i >= args.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals]) * 1e3)
    one, _, _ = cpu_arm(1, nseg, seed=99)        # the faithful figure: one decode thread per file, files sequential (processor.rs:31, lib.rs:694)
    sample = (f"{cores} worker processes x {nseg} windows (10 min of C2 audio each) per step; {CPU_SAMPLE_NOTE}; "
              "value = all cores (N cooperating birda processes), single_thread_value = one process")
    line = {"impl": "reference", "metric": "audio-hours/sec", "value": value, "unit": "audio-h/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference CPU algorithm restated (oracle); the Rust binary cannot be built here"},
            "cpu_baseline": {"value": value, "unit": "audio-h/s", "cores": cores, "kind": "port", "sample": sample,
                             "single_thread_value": one},
            "e2e": {"value": value, "unit": "audio-h/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Best effort: run this rank (and allocate its pinned staging) on the NUMA node its GPU hangs off, so the
    H2D stream does not cross the socket interconnect.  Returns a note for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b_ = part.partition("-")
            ids += list(range(int(a), int(b_ or a) + 1))
        allowed = sorted(set(ids) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa: rank bound to node {node} ({len(allowed)} cpus)"
        return f"numa: node {node} has no allowed cpus"
    except Exception as e:                                   # not fatal: containers often hide /sys topology
        return f"numa: unavailable ({type(e).__name__})"


# ------------------------------------------------------------------------------------------ GPU arm
def side_kernels(torch, b, ctx, device, timed, steps):
    """The path's other kernels on their own workloads (rank 0, outside the headline step): K1 on the bat config C4
    (no resampling: the HBM-bound kernel) and K5, the optional spectrogram prefix (tcgen05 mel projection)."""
    out = {}
    steps = max(3, min(steps, 10))
    # K1: C4, 20 min of 256 kHz mono s16 -> 2846 windows of 144000 samples, hop 108000
    n = 20 * 60 * 256_000
    pcm = (torch.randn(n, device=device) * 3000).to(torch.int16)
    plan = b.FrontEndPlan(ctx, 256_000, 1, b.FMT_S16, 256_000, 144_000, 36_000)
    r = plan.run(pcm, pad_to_batch=1, want_tables=False)
    ms = timed(lambda: plan.run(pcm, pad_to_batch=1, want_tables=False), steps) / steps
    algo = n * 2 + r.nseg * 144_000 * 4
    peak, _ = measured_peaks()
    out["k1_pack_c4"] = {"workload": "C4 bat: 20 min 256 kHz mono s16, 2846 windows, overlap 36000", "ms": ms,
                         "algorithmic_bytes": algo, "achieved_gbs": algo / ms / 1e6, "frac_of_hbm_peak": algo / ms / 1e6 / peak}
    plan.close(); del pcm
    # K5: BirdNET-v2.4-like low-band spectrogram of 512 packed windows (n_fft 2048, hop 278, 511 frames, 96 mels 0-3 kHz)
    rows, samples, n_fft, hop, n_frames, n_mels = 512, SEG, 2048, 278, 511, 96
    bins = n_fft // 2 + 1
    mel = lambda f: 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    m = mel(np.linspace(0.0, TGT_RATE / 2.0, bins)); edges = np.linspace(mel(0.0), mel(3000.0), n_mels + 2)
    mw = np.zeros((n_mels, bins), dtype=np.float32)
    for i in range(n_mels):
        mw[i] = np.maximum(0.0, np.minimum((m - edges[i]) / (edges[i + 1] - edges[i]), (edges[i + 2] - m) / (edges[i + 2] - edges[i + 1])))
    mw[:, 0] = 0
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)
    x = torch.randn((rows, samples), device=device) * 0.1
    o = torch.empty((rows, n_mels, n_frames), device=device)
    ms_obj = b.MelSpec(ctx, n_fft, hop, n_frames, win, mw)
    _, nb, kpad = ms_obj.info()
    ms_obj.run(x.data_ptr(), rows, samples, o.data_ptr())
    ms = timed(lambda: ms_obj.run(x.data_ptr(), rows, samples, o.data_ptr()), steps) / steps
    out["k5_melspec_low_band"] = {"workload": f"{rows} windows x {n_frames} frames, n_fft {n_fft}, hop {hop}, {n_mels} mels over {nb} bins (K {kpad})",
                                  "ms": ms, "audio_hours_per_s_at_hop_1.5s": rows * 1.5 / 3600.0 / (ms / 1e3),
                                  "gemm_gflop_3xtf32": 3 * 2.0 * rows * n_frames * kpad * n_mels / 1e9,
                                  "note": "STFT (CUDA cores) dominates; tensor-pipe share of the GEMM kernel is in profiles/"}
    ms_obj.close()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import birda_b200 as b

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    pcm, scores, mask = make_workload_gpu(torch, device, seed=2 + rank)
    stream = torch.cuda.current_stream()
    ctx = b.Context(local, stream=stream.cuda_stream)
    plan = b.FrontEndPlan(ctx, SRC_RATE, CHANNELS, b.FMT_S16, TGT_RATE, SEG, OVL)
    cfg = b.PostConfig(activation=b.ACT_SIGMOID, min_confidence=0.1, top_k=5, range_threshold=0.01,
                       keep_unmatched=True, rerank=False)
    rows = 2432
    d_idx = torch.empty((rows, 5), dtype=torch.int32, device=device)
    d_conf = torch.empty((rows, 5), dtype=torch.float32, device=device)
    d_cnt = torch.empty((rows,), dtype=torch.int32, device=device)
    h_idx = torch.empty((rows, 5), dtype=torch.int32).pin_memory()
    h_conf = torch.empty((rows, 5), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((rows,), dtype=torch.int32).pin_memory()
    h_pcm = torch.empty(pcm.shape, dtype=torch.int16).pin_memory()
    h_pcm.copy_(pcm)
    torch.cuda.synchronize()

    def step_resident():
        r = plan.run(pcm, pad_to_batch=BATCH, want_tables=False)
        ctx.post_run_device(scores.data_ptr(), rows, CLASSES, r.nseg, cfg, mask.data_ptr(), None,
                            d_idx.data_ptr(), d_conf.data_ptr(), d_cnt.data_ptr())
        return r

    def step_e2e():
        r = plan.run(h_pcm.numpy(), pad_to_batch=BATCH, want_tables=False)
        ctx.post_run_device(scores.data_ptr(), rows, CLASSES, r.nseg, cfg, mask.data_ptr(), None,
                            d_idx.data_ptr(), d_conf.data_ptr(), d_cnt.data_ptr())
        h_idx.copy_(d_idx, non_blocking=True); h_conf.copy_(d_conf, non_blocking=True); h_cnt.copy_(d_cnt, non_blocking=True)
        return r

    # front end + (stand-in) inference + post: the library's stand-in classifier (NOT BirdNET: 48 band energies times a
    # fixed matrix, the real [B, 144000] -> [B, 6522] contract) over the 38 batches the front end packed, its logits
    # through the post step batch by batch — the per-file loop of the product (csrc/pipeline.cpp) on resident PCM
    standin = b.StandIn(local, SEG, CLASSES, BATCH, seed=11, stream=stream.cuda_stream)
    import ctypes as C

    from birda_b200 import _lib

    def step_with_standin():
        r = plan.run(pcm, pad_to_batch=BATCH, want_tables=False)
        base = r.device_ptr
        ds, nc = C.c_void_p(), C.c_uint32()
        for first in range(0, r.rows, BATCH):
            rc = _lib.lib.bb_standin_classify(standin.handle, C.c_void_p(base + first * SEG * 4), BATCH, SEG, C.byref(ds), C.byref(nc))
            assert rc == 0
            valid = min(BATCH, r.nseg - first)
            ctx.post_run_device(ds.value, BATCH, CLASSES, valid, cfg, mask.data_ptr(), None,
                                d_idx.data_ptr() + first * 20, d_conf.data_ptr() + first * 20, d_cnt.data_ptr() + first * 4)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        r = step_resident()
    torch.cuda.synchronize()
    assert r.nseg == 2400 and r.rows == rows, (r.nseg, r.rows)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.kernel_launches
    ms_total = timed(step_resident, args.steps)
    launches = ctx.kernel_launches - l0
    # dominant kernel alone (K2 front end) for the roofline, same stream, same events
    ms_front = timed(lambda: plan.run(pcm, pad_to_batch=BATCH, want_tables=False), args.steps) / args.steps
    # the post step alone, on three score buffers in rotation (188 MB > the 126 MB L2: every pass reads DRAM, as in the real
    # step where K2 streams 2 GB between two post steps)
    score_ring = [scores, scores.clone(), scores.clone()]
    ring_pos = [0]

    def post_alone():
        sc = score_ring[ring_pos[0] % 3]; ring_pos[0] += 1
        ctx.post_run_device(sc.data_ptr(), rows, CLASSES, 2400, cfg, mask.data_ptr(), None,
                            d_idx.data_ptr(), d_conf.data_ptr(), d_cnt.data_ptr())

    for _ in range(3):
        post_alone()
    ms_post = timed(post_alone, args.steps) / args.steps
    del score_ring[1:]
    step_with_standin()
    sl0, cl0 = standin.launches, ctx.kernel_launches
    ms_standin = timed(step_with_standin, args.steps)
    launches_standin = (standin.launches - sl0) + (ctx.kernel_launches - cl0)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # the same bytes as one plain pinned -> device copy: the PCIe floor of the end-to-end step on this box
    ms_h2d_only = timed(lambda: pcm.copy_(h_pcm, non_blocking=True), args.steps) / args.steps
    def timed_local(fn, steps):                 # rank-local (no collective): side kernels run on rank 0 only
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return float(e0.elapsed_time(e1))

    extras = side_kernels(torch, b, ctx, device, timed_local, args.steps) if rank == 0 else None
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        hours = SECONDS / 3600.0
        value = args.steps * hours * world / (ms_total / 1e3)
        e2e = args.steps * hours * world / (ms_e2e / 1e3)
        peak, peak_src = measured_peaks()
        ach = ALGO_BYTES_FRONT / (ms_front / 1e3) / 1e9
        kernel = plan.describe()
        traffic, smem_pct, prof_src = k2_profile(kernel)
        h2d_bytes = int(h_pcm.numel() * 2)
        cores = len(os.sched_getaffinity(0)) or 1        # the threads this process may actually use
        cpu_v, cpu_wall, _ = cpu_arm(cores, 400, seed=7)      # 10 min of audio per core, a few seconds
        cpu_one, _, _ = cpu_arm(1, 400, seed=8)
        line = {
            "metric": "audio-hours/sec", "value": value, "unit": "audio-h/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sharding": f"file-sharded, {world} x 1 audio-hour file per step, no collective",
                       "l2": "inputs (635 MB PCM) and outputs (1.4 GB) per step exceed the 126 MB L2",
                       "model_forward": "not executed (not on the path; ONNX Runtime absent) - scores synthetic, resident",
                       "ms_front_end": ms_front, "ms_post": ms_post, "host": numa_note,
                       "front_end_plus_standin_inference": {
                           "value": args.steps * hours * world / (ms_standin / 1e3), "unit": "audio-h/s", "ms_per_step": ms_standin / args.steps,
                           "gpu_launches": int(launches_standin),
                           "note": "front end + the library's STAND-IN classifier (not BirdNET; same [64,144000]->[64,6522] contract, "
                                   "38 batches) + post step per batch, inputs resident"},
                       "other_kernels": extras},
            "e2e": {"value": e2e, "unit": "audio-h/s", "h2d_bytes_per_step": int(h_pcm.numel() * 2),
                    "d2h_bytes_per_step": int(rows * 5 * 8 + rows * 4), "ms_per_step": ms_e2e / args.steps,
                    "h2d_copy_alone_ms": ms_h2d_only, "frac_of_pcie_floor": ms_h2d_only / (ms_e2e / args.steps),
                    # the same bytes as ONE plain pinned->device copy per rank, all ranks at once (max over ranks): what the
                    # host's DMA path gives this box at this N; e2e cannot beat it
                    "host_dma_floor": {"gbs_per_gpu": h2d_bytes / ms_h2d_only / 1e6, "aggregate_gbs": world * h2d_bytes / ms_h2d_only / 1e6,
                                       "audio_h_per_s": hours * world / (ms_h2d_only / 1e3)}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "algorithmic_bytes": ALGO_BYTES_FRONT,
                         "peak_source": peak_src,
                         "note": "K2 is FP32/shared-memory bound, not HBM bound (SURVEY 7.3 item 2); algorithmic flops are the reference "
                                 "blocking's (SURVEY 8d): the kernel's own blocking does about a third less transform work",
                         "shared_memory_pipe_pct_ncu": smem_pct, "profile_source": prof_src,
                         "fp32": {"achieved_tflops": ALGO_FLOPS_FRONT / (ms_front / 1e3) / 1e12, "peak_tflops": 74.5,
                                  "frac": ALGO_FLOPS_FRONT / (ms_front / 1e3) / 1e12 / 74.5}},
            "cpu_baseline": {"value": cpu_v, "unit": "audio-h/s", "cores": cores, "kind": "port", "single_thread_value": cpu_one,
                             "sample": f"{cores} worker processes x 400 windows (10 min of C2 audio each); {CPU_SAMPLE_NOTE}; "
                                       "single_thread_value = one process (the reference decodes on one thread per file)"},
            "clocks": clocks,
        }
        print(json.dumps(line))
    standin.close(); plan.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ C5: directory batch
C5_KINDS = [(16_000, 1), (22_050, 2), (32_000, 1), (44_100, 2), (48_000, 1), (96_000, 2)]     # rates and channel counts cycling
C5_WORKLOAD = ("C5: directory of {n} synthetic 10-min s16 WAV files at mixed rates (16/22.05/32/44.1/48/96 kHz, mono/stereo) in "
               "/dev/shm, BirdNET v2.4 windows (3 s, overlap 0, batch 64), stand-in classifier, range mask, through bb_pool_process_wavs")


def c5_make_files(n_files: int, seconds: float, directory: str, copies: bool = True):
    """12 master files (6 kinds x 2 seeds: one synthetic minute tiled, perturbed) copied to n_files files (own pages in
    the page cache each, as a real directory has; `copies=False` hard-links them instead: less RAM, but readers of the
    same inode then contend for its pages)."""
    import shutil
    from birda_b200.synth import synth_pcm, write_wav
    os.makedirs(directory, exist_ok=True)
    masters = {}
    for k, (sr, ch) in enumerate(C5_KINDS):
        for v in range(2):
            path = os.path.join(directory, f"master_{k}_{v}.wav")
            if not os.path.exists(path):
                base = synth_pcm(1000 + 2 * k + v, 60.0, sr, ch).reshape(-1, ch)
                reps = int(np.ceil(seconds / 60.0))
                pcm = np.tile(base, (reps, 1))[: int(seconds * sr)].copy()
                pcm[:: 9973, 0] += 17
                write_wav(path + ".tmp", pcm.reshape(-1), sr, ch)
                os.replace(path + ".tmp", path)
            masters[(k, v)] = path
    paths = []
    for i in range(n_files):
        p = os.path.join(directory, f"file_{i:04d}.wav")
        if not os.path.exists(p):
            if copies:
                shutil.copyfile(masters[(i % 6, (i // 6) % 2)], p + ".tmp"); os.replace(p + ".tmp", p)
            else:
                os.link(masters[(i % 6, (i // 6) % 2)], p)
        paths.append(p)
    return paths


def run_c5(args):
    """BASELINE config 5 through the product's multi-GPU path: files sharded over the ranks (longest first), each rank
    runs the library's file pool (csrc/pool.cpp) with `--workers` contexts on its GPU and the library's stand-in
    classifier as a native callback (no Python in the loop).  Strong scaling: the directory is fixed, ranks share it.
    Timed with the host clock between barriers (max over ranks): the path is host-driven — file reads, worker threads,
    H2D copies — and no device event sees all of it."""
    import torch
    import torch.distributed as dist

    import birda_b200 as b
    from birda_b200.pipeline import NativePool, ProcessingConfig
    from birda_b200.shard import shard_files

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    seconds, directory = 600.0, "/dev/shm/birda_b200_c5"
    if rank == 0:
        paths = c5_make_files(args.files, seconds, directory, copies=not args.hardlinks)
    if world > 1:
        dist.barrier()
    paths = [os.path.join(directory, f"file_{i:04d}.wav") for i in range(args.files)]
    # longest-processing-time-first by BYTES: the files are equally long, what a rank pays for is their PCM (16 kHz mono
    # to 96 kHz stereo: 19 - 230 MB) — by duration alone a round robin would hand one rank all the stereo files
    mine = shard_files([float(os.path.getsize(p)) for p in paths], world)[rank]
    my_paths = [paths[i] for i in mine]
    my_bytes = sum(os.path.getsize(p) for p in my_paths)
    g = torch.Generator(device=device); g.manual_seed(5)
    mask = torch.rand(CLASSES, generator=g, device=device) ** 2
    mask[torch.randperm(CLASSES, generator=g, device=device)[:305]] = float("nan")
    cfg = ProcessingConfig(target_rate=TGT_RATE, segment_duration=3.0, overlap=0.0, batch_size=BATCH, min_confidence=0.1,
                           d_mask=mask.data_ptr(), range_threshold=0.01, keep_unmatched=True, rerank=False)
    W = max(1, args.workers)
    standins = [b.StandIn(local, SEG, CLASSES, BATCH, seed=11) for _ in range(W)]
    pool = NativePool([local] * W, [cfg] * W, standins)
    if not args.pool_sync:
        pool.bind_standins_to_worker_streams()      # the classifier queues on its worker's stream: one wait per file

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    res = None
    for _ in range(max(args.warmup, 1)):
        res = pool.process_wavs_counts(my_paths)
    if rank == 0:
        sampler.start()
    l0 = pool.kernel_launches + sum(s.launches for s in standins)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = pool.process_wavs_counts(my_paths)       # detections stay C arrays: the library is timed, not a Python conversion
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    dt = float(t.item())
    launches = pool.kernel_launches + sum(s.launches for s in standins) - l0
    clocks = sampler.stop() if rank == 0 else None
    nseg = sum(r[0] for r in res); ndet = sum(r[1] for r in res)
    assert all(r[0] == 200 for r in res), "every 10-min file is 200 windows at overlap 0"
    if rank == 0:
        hours = args.files * seconds / 3600.0
        value = args.steps * hours / dt
        cores = len(os.sched_getaffinity(0)) or 1
        line = {"metric": "audio-hours/sec", "value": value, "unit": "audio-h/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": C5_WORKLOAD.format(n=args.files), "files": args.files, "workers_per_gpu": W, "classifier_stream": "own, waits per batch" if args.pool_sync else "worker's stream, one wait per file",
                           "sharding": f"files sharded longest-first over {world} rank(s), no collective", "host": numa_note,
                           "timer": "host wall clock between barriers, max over ranks (host-driven path: file reads + worker threads + copies)",
                           "l2": "every file's PCM (19-230 MB) is read from page cache, copied H2D and packed to 115 MB of windows: > L2 per file",
                           "rank0_segments": nseg, "rank0_detections": ndet, "host_cores": cores},
                "e2e": {"value": value, "unit": "audio-h/s", "h2d_bytes_per_step": int(my_bytes), "d2h_bytes_per_step": int(nseg * 44),
                        "note": "this workload IS end to end: WAV files in page cache -> detections on the host"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line))
    pool.close()
    for s_ in standins:
        s_.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c2: the headline step (default); c5: directory batch through the file pool")
    ap.add_argument("--files", type=int, default=1000, help="c5: files in the directory")
    ap.add_argument("--workers", type=int, default=6, help="c5: pool workers (contexts) per GPU")
    ap.add_argument("--pool-sync", action="store_true", help="c5: classifier on its own stream, pool waits around every batch (the default contract of bb_pool)")
    ap.add_argument("--hardlinks", action="store_true", help="c5: hard-link the 12 master files instead of copying them (84 GB of page cache at 1000 files otherwise)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        if args.steps == 20:
            args.steps = 2
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
