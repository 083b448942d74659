#!/usr/bin/env python3
"""Turn one tools/gpu_profile_round.sh run (gpurun_out/<TAG>_*) into the tracked summaries under profiles/.
usage: python tools/collect_profiles.py TAG [PREFIX]     (PREFIX of the files written under profiles/, default r02_final)"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
PREFIX = sys.argv[2] if len(sys.argv) > 2 else "r02_final"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
NOTES = {
    "k2_c2_1h": "K2 resample_plan2_kernel (two-stream CT plan, own blocking 4116->4480-point transforms, 640 threads), full C2 audio-hour, one launch",
    "k2_c3_1h": "K2 resample_plan2_kernel (two-stream CT plan 1026/684, 384 threads), full C3 audio-hour (Perch 48->32 kHz), one launch",
    "k1_c4": "K1 pack kernel, C4 bat 256 kHz mono s16, 20 min of audio (contiguous warp stores)",
    "k3_c2": "K3 post kernel, 2400 x 6522 scores, 64-thread CTAs",
    "k5a_low": "K5a stft_power2_kernel (two-stream, radices 16 8 8, padded address map), 297 windows x 511 frames, n_fft 2048, low band",
    "k5b_low": "K5b mel_gemm_kernel (tcgen05 kind::tf32 x3), 297 windows x 511 frames, low band K=128, 96 mels",
    "k5b_full": "K5b mel_gemm_kernel (tcgen05 kind::tf32 x3), 297 windows x 511 frames, full band K=1024, 128 mels",
}
for k, note in NOTES.items():
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if os.path.exists(rep):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, os.path.join(P, f"{PREFIX}_{k}.txt"),
                        PREFIX + " " + note], stdout=subprocess.DEVNULL, check=True)
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10]
ci = {k: i for i, k in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "")
    v, u = float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]]
    agg.setdefault(name, []).append(v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3))
tot = sum(sum(v) for v in agg.values())
lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'resample|pack_kernel|post_kernel|stft_power|mel_gemm' python bench.py --steps 3 --warmup 3",
         "# (cold-cache, serialised: compare SHARES; covers warm-up, resident, K2-only, K3-only, e2e (K2 in 16 pieces) and the side-kernel loops of bench.py)"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k[:70]:70s} launches={len(v):4d} mean_us={sum(v) / len(v):9.1f} total_ms={sum(v) / 1e3:8.2f} share={sum(v) * 100 / tot:5.1f}%")
b = json.loads(open(os.path.join(G, f"{tag}_bench.json")).read().strip().splitlines()[-1])
c = b["config"]
lines.append(f"# live (bench.py, CUDA events, same box): ms_front_end {c['ms_front_end']:.3f}, ms_post {c['ms_post']:.3f} -> "
             f"K2 share of the headline step {c['ms_front_end'] * 100 / (c['ms_front_end'] + c['ms_post']):.1f} %")
open(os.path.join(P, f"{PREFIX}_launches_bench_summary.txt"), "w").write("\n".join(lines) + "\n")
open(os.path.join(P, f"{PREFIX}_launches_bench.csv"), "w").write(open(os.path.join(G, f"{tag}_launches.csv")).read())
open(os.path.join(P, f"{PREFIX}_bench_line.json"), "w").write(json.dumps(b) + "\n")
open(os.path.join(P, f"{PREFIX}_bench_reference_line.json"), "w").write(open(os.path.join(G, f"{tag}_bench_ref.json")).read().strip().splitlines()[-1] + "\n")
# DRAM traffic of the dominant kernel for bench.py's roofline.traffic
txt = open(os.path.join(P, f"{PREFIX}_k2_c2_1h.txt")).read()
def val(name):
    m = re.search(name + r"\s+([0-9.]+) (\w+)", txt)
    x, u = float(m.group(1)), m.group(2)
    return int(x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
m = re.search(r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed\s+([0-9.]+)", txt)
# the kernel as the library names it on the box the capture was taken on (bb_plan_describe): bench.py quotes these
# numbers only when its own run launches the same kernel
kernel = open(os.path.join(G, f"{tag}_k2_describe.txt")).read().strip()
json.dump({"kernel": kernel, "workload": "C2 audio-hour (2400 windows + 32 padding rows), one launch",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "algorithmic_bytes": 2017440000,
           "smem_pipe_pct": float(m.group(1)) if m else None,
           "source": f"profiles/{PREFIX}_k2_c2_1h.txt (ncu --set full)"}, open(os.path.join(P, "k2_traffic.json"), "w"), indent=1)
print("\n".join(lines[2:]))
print("value", b["value"], "e2e", b["e2e"]["value"], "roofline.frac", b["roofline"]["frac"], "k1", c["other_kernels"]["k1_pack_c4"]["frac_of_hbm_peak"])
