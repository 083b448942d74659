#!/usr/bin/env python3
"""Per-phase view of a kernel from an .ncu-rep: the SASS listing is cut at every BAR.SYNC (= stage boundary of the K2
thread groups) and shared-memory wavefronts, excess (bank-conflict) wavefronts, executed instructions and stall samples are
summed per segment.  usage: python tools/ncu_segments.py x.ncu-rep [min_share_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[col[k]])
    except (ValueError, IndexError):
        return 0.0
segs = []
cur = dict(first=None, n=0, inst=0.0, wf=0.0, ex=0.0, samples=0.0, lds=0, sts=0, bar=0.0, ssb=0.0, ops={})
tot = dict(inst=0.0, wf=0.0, ex=0.0, samples=0.0)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    if cur["first"] is None:
        cur["first"] = r[col["Address"]]
    cur["n"] += 1
    i, w, e, s = num(r, "Instructions Executed"), num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Excessive"), num(r, "# Samples")
    cur["inst"] += i; cur["wf"] += w; cur["ex"] += e; cur["samples"] += s
    cur["bar"] += num(r, "stall_barrier"); cur["ssb"] += num(r, "stall_short_sb")
    op = src.split()[0] if src.split() else ""
    if op.startswith("@"):
        op = src.split()[1] if len(src.split()) > 1 else op
    key = op.split(".")[0]
    cur["ops"][key] = cur["ops"].get(key, 0) + i
    tot["inst"] += i; tot["wf"] += w; tot["ex"] += e; tot["samples"] += s
    if "BAR.SYNC" in src or "EXIT" in src:
        segs.append(cur)
        cur = dict(first=None, n=0, inst=0.0, wf=0.0, ex=0.0, samples=0.0, lds=0, sts=0, bar=0.0, ssb=0.0, ops={})
if cur["n"]:
    segs.append(cur)
print(f"total: inst {tot['inst']:.3e}  smem wavefronts {tot['wf']:.3e}  excess {tot['ex']:.3e} ({100*tot['ex']/max(tot['wf'],1):.1f} %)  samples {tot['samples']:.0f}")
print(f"{'seg':>3} {'addr':>8} {'sass':>5} {'inst%':>6} {'wf%':>6} {'excess/wf%':>10} {'samples%':>8} {'bar%':>6} {'ssb%':>6}  top ops")
for k, s in enumerate(segs):
    if s["samples"] < tot["samples"] * min_share / 100 and s["wf"] < tot["wf"] * min_share / 100:
        continue
    ops = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"{k:3d} {s['first'][-6:]:>8} {s['n']:5d} {100*s['inst']/tot['inst']:6.1f} {100*s['wf']/max(tot['wf'],1):6.1f} "
          f"{100*s['ex']/max(s['wf'],1):10.1f} {100*s['samples']/max(tot['samples'],1):8.1f} {100*s['bar']/max(s['samples'],1):6.1f} {100*s['ssb']/max(s['samples'],1):6.1f}  "
          + " ".join(f"{o}:{100*v/max(s['inst'],1):.0f}" for o, v in ops))
