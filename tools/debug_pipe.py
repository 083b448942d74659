import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import birda_b200 as b
from birda_b200.synth import synth_pcm
from oracle import frontend as ofe, post as opost, rules as orules
sys.path.insert(0, "tests")
from test_gpu_pipeline import StandInClassifier
ctx = b.Context(0)
pcm = synth_pcm(11, 61.0, 44100, 2)
plan = b.FrontEndPlan(ctx, 44100, 2, b.FMT_S16, 48000, 144000, 72000)
segs = plan.run(pcm, pad_to_batch=16); ctx.sync()
got = segs.torch().cpu().numpy()
ref = ofe.decode_and_stream(pcm, 2, 44100, 48000, 144000, 72000)
rms = np.sqrt(np.mean(ref.segments.astype(np.float64)**2, axis=1))
err = np.abs(got[:segs.nseg] - ref.segments).max(axis=1) / np.maximum(rms, 1e-30)
print("nseg", segs.nseg, "rows", segs.rows, "max rel err per row:", np.round(err, 8))
clf = StandInClassifier(144000, 6522)
s1 = clf(torch.from_numpy(got[:16]).cuda()).cpu().numpy(); s2 = clf(torch.from_numpy(ref.segments[:16]).cuda()).cpu().numpy()
print("score diff max", np.abs(s1 - s2).max(), "score range", s1.min(), s1.max())
c1 = 1/(1+np.exp(-s1)); print("n conf>=0.1 per row", (c1 >= 0.1).sum(axis=1))
