"""Error of the coarse-map routes against the raw-map evaluation, across envmaps and roughness (diagnostic, GPU)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
dev = "cuda:0"
def run(env, z, v, coarse):
    os.environ["DRM_RENDER_COARSE"] = "1" if coarse else "0"
    return render_batch(env, z, v, res=64, footprint_S=1)
worst = {}
for seed in range(1000, 1012):
    env = synthetic_envmap(1000, 2000, seed, device=dev, as_numpy=False)[None]
    for rough in (0.5, 0.55, 0.62, 0.7, 0.72, 0.8, 1.0):
        for metal in (1.0, 0.0):
            z = torch.tensor([[metal, 0.9, 0.7, 0.5, rough, 0.8]]); v = torch.tensor([[0.4, 0.0, 1.0]])
            a = run(env, z, v, True); b = run(env, z, v, False)
            err = ((a - b).norm() / b.norm()).item()
            key = (rough, metal)
            worst[key] = max(worst.get(key, 0), err)
for k in sorted(worst): print(k, f"{worst[k]:.2e}")
