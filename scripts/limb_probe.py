"""Limb window (n.v -> 0) of a forced S = 16 render against the fp64 oracle: hierarchy, single level, stage toggles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drmnet_b200.synth import synthetic_envmap
from drmnet_b200.renderer import render_batch
from oracle.render_oracle import render_oracle

z = [float(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1.0, 0.95, 0.6, 0.3, 0.3, 1.0]
view = [0.0, 0.0, 1.1]
env = synthetic_envmap(500, 1000, seed=1004)
envd = torch.from_numpy(env).cuda()[None]
win = (62, 66, 124, 128)
ref = render_oracle(env, z, view, 128, S=16, window=win)[win[0]:win[1], win[2]:win[3]]
def run(off):
    for k in list(os.environ):
        if k.startswith("DRM_RENDER_"): os.environ.pop(k)
    for k in off: os.environ["DRM_RENDER_" + k] = "0"
    o = render_batch(envd, torch.tensor([z]), torch.tensor([view]), res=128, footprint_S=16, channel_first=False)[0].cpu().numpy()
    return o[win[0]:win[1], win[2]:win[3]]
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
flat = run(["LEVELS", "COARSE"])
print("flat vs oracle %.2e" % rel(flat, ref))
for off in ([], ["NEAR"], ["VIEW_AVG"], ["FAR_COARSE"], ["FAR_COARSE4"], ["COARSE"], ["DIFF_CORR"]):
    o = run(off)
    print("off %-12s vs oracle %.2e  vs flat %.2e" % (",".join(off), rel(o, ref), rel(o, flat)))
print("per-cell (hier-ref)/ref, channel 0:\n", ((run([]) - ref) / ref)[..., 0])
