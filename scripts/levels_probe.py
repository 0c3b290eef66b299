"""Footprint-hierarchy error and time against the single-level brute force of the same canonical sum (diagnostic, GPU).
The scales given on the command line are values of DRM_RENDER_NEAR_SCALE (per-cell near-field thresholds)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch, auto_footprint
from drmnet_b200.synth import synthetic_envmap

dev = "cuda:0"
def run(env, z, v, S, levels, scale=1.0):
    os.environ["DRM_RENDER_LEVELS"] = "1" if levels else "0"
    os.environ["DRM_RENDER_NEAR_SCALE"] = str(scale)
    torch.cuda.synchronize(); t = time.time()
    o = render_batch(env, z, v, res=128, footprint_S=S)
    torch.cuda.synchronize()
    return o, time.time() - t

scales = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1.0", "0.6"])]
for seed in (1004,):
    env = synthetic_envmap(1000, 2000, seed, device=dev, as_numpy=False)[None]
    for rough, metal in ((0.0, 1.0), (0.08, 1.0), (0.11, 0.5), (0.15, 1.0), (0.18, 0.2), (0.22, 0.0), (0.3, 1.0), (0.4, 1.0)):
        z = torch.tensor([[metal, 0.9, 0.7, 0.5, rough, 0.8]])
        v = torch.tensor([[0.4, 0.0, 1.0]])
        S = auto_footprint(rough, 128, 1.25 * 3.141592653589793 / 1000)
        ref, t0 = run(env, z, v, S, False)
        line = f"seed {seed} rough {rough:4.2f} metal {metal} S={S:2d} brute {t0*1e3:8.1f} ms |"
        for sc in scales:
            o, t1 = run(env, z, v, S, True, sc)
            err = ((o - ref).norm() / ref.norm()).item()
            line += f" scale {sc}: {t1*1e3:7.1f} ms err {err:.2e} |"
        print(line, flush=True)
