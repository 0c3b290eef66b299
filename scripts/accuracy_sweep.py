"""Accelerated render against the un-accelerated sum (every DRM_RENDER_* stage off) on random renders (GPU).

usage: accuracy_sweep.py ENV_SEED0 Z_SEED0 RENDERS    -> one line per render: roughness, metallic, footprint, rel-L2,
       worst cell error relative to the image maximum
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drmnet_b200 import synth
from drmnet_b200.renderer import render_batch, auto_footprint

STAGES = ("DRM_RENDER_COARSE", "DRM_RENDER_LEVELS", "DRM_RENDER_NEAR", "DRM_RENDER_FAR_COARSE", "DRM_RENDER_FAR_COARSE4",
          "DRM_RENDER_UNIFY", "DRM_RENDER_DIFF_CORR", "DRM_RENDER_VIEW_AVG")


def main():
    dev = "cuda:0"
    He, We, B = 1000, 2000, 8
    e0, z0, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    env = torch.stack([synth.synthetic_envmap(He, We, e0 + i, device=dev) for i in range(B)])
    z = torch.stack([synth.sample_brdf(z0 + i) for i in range(R)]).to(dev)
    view = torch.stack([synth.sample_view(z0 + i) for i in range(R)]).to(dev)
    idx = (torch.arange(R, device=dev) % B).int()
    fast = render_batch(env, z, view, env_index=idx, res=128, footprint_S=None)
    for k in STAGES:
        os.environ[k] = "0"
    full = render_batch(env, z, view, env_index=idx, res=128, footprint_S=None)
    torch.cuda.synchronize()
    rel = torch.linalg.norm((fast - full).flatten(1), dim=1) / torch.linalg.norm(full.flatten(1), dim=1)
    cell = (fast - full).abs().flatten(1).amax(1) / full.flatten(1).amax(1)
    print(f"# envmap seeds {e0}..{e0 + B - 1}, BRDF/view seeds {z0}..{z0 + R - 1}; max rel-L2 {rel.max().item():.2e}, "
          f"max cell error / image max {cell.max().item():.2e}")
    for i in range(R):
        print(f"{i:3d} rough {z[i, 4].item():.3f} metal {z[i, 0].item():.2f} S {auto_footprint(float(z[i, 4]), 128):2d} "
              f"rel-L2 {rel[i].item():.2e} cell/max {cell[i].item():.2e}")


if __name__ == "__main__":
    main()
