"""Local (per-pixel) effect of gathering the far field of sharp lobes from the 2x2 coarse map (diagnostic, GPU)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import Z0, synthetic_envmap
dev = "cuda:0"
for seed in (1001, 1004, 1007, 1010):
    env = synthetic_envmap(1000, 2000, seed, device=dev, as_numpy=False)[None]
    for z, S in ((list(Z0), 16), ([1.0, 0.9, 0.8, 0.7, 0.09, 1.0], 8)):
        zt = torch.tensor([z]); v = torch.tensor([[0.4, 0.0, 1.0]])
        os.environ["DRM_RENDER_FAR_COARSE"] = "1"; a = render_batch(env, zt, v, res=128, footprint_S=S)
        os.environ["DRM_RENDER_FAR_COARSE"] = "0"; b = render_batch(env, zt, v, res=128, footprint_S=S)
        rel = ((a - b).abs() / b.abs().clamp_min(1e-6))
        print(f"seed {seed} S {S}: global rel-L2 {((a-b).norm()/b.norm()).item():.2e}  max per-pixel rel {rel.max().item():.2e}  99.9th pct {rel.flatten().kthvalue(int(rel.numel()*0.999)).values.item():.2e}")
