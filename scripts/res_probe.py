"""Hierarchical render against the single-level kernel at refmap resolutions other than 128 (BASELINE config[4] sweeps
64^2 .. 256^2): whole-image relative L2 and the worst cell relative to its own value.  Development tool, GPU."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import Z0, sample_brdf, schedule_point, synthetic_envmap

def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def local(a, b):
    return float((np.abs(a - b).max(-1) / (np.abs(b).max(-1) + 1e-2 * np.median(b))).max())

Z = {"z0_mirror": list(Z0), "near_mirror": schedule_point(sample_brdf(11), 0.1)[2].tolist(),
     "glossy_metal": [1.0, 0.95, 0.6, 0.3, 0.3, 1.0], "mixed": [0.4, 0.3, 0.8, 0.6, 0.45, 0.8],
     "rough_dielectric": [0.0, 0.9, 0.5, 0.2, 0.7, 0.5], "gloss15": [0.2, 0.7, 0.7, 0.9, 0.15, 0.8],
     # the wide end of the 2x2 footprint at res 128 (cell / alpha just above 0.1), with and without a diffuse lobe
     "edge_s2_dielectric": [0.0, 0.9, 0.5, 0.2, 0.48, 0.5], "edge_s2_metal": [1.0, 0.9, 0.5, 0.2, 0.48, 0.5]}
from drmnet_b200 import _lib
OPT = _lib.default_render_options()
for a in [x for x in sys.argv[1:] if ":" in x]:  # option overrides, e.g. horizon_finest:0.03
    k, v = a.split(":")
    setattr(OPT, k, type(getattr(OPT, k))(float(v)))
ONLY = [x for x in sys.argv[1:] if not x.isdigit() and ":" not in x]
if ONLY:
    Z = {k: v for k, v in Z.items() if k in ONLY}
RES = [int(x) for x in sys.argv[1:] if x.isdigit()] or [64, 256]
for He, We in ((500, 1000), (1000, 2000)):
    env = synthetic_envmap(He, We, seed=1004, device="cuda")[None]
    for res in RES if He == 500 else [r for r in RES if r <= 64]:
        for name, z in Z.items():
            for v in ([0.644, 0.0, 0.765], [-0.5, 0.3, -0.8]):
                zz, vv = torch.tensor([z]), torch.tensor([v])
                tree = render_batch(env, zz, vv, res=res, footprint_S=None, channel_first=False, check_status=True, options=OPT)[0].cpu().numpy()
                S = int(render_batch.last_status[7]) if False else None
                # the footprint the device chose is not reported: use the host rule for the single-level kernel
                from drmnet_b200.renderer import auto_footprint, default_alpha_min
                S = auto_footprint(float(np.clip(z[4], 0, 1)), res, default_alpha_min(He))
                flat = render_batch(env, zz, vv, res=res, footprint_S=S, channel_first=False, flat=True)[0].cpu().numpy()
                tree_s = render_batch(env, zz, vv, res=res, footprint_S=S, channel_first=False, options=OPT)[0].cpu().numpy()
                print(f"{He}x{We} res {res:3d} {name:16s} view {v} S={S:2d}: rel-L2 {rel_l2(tree_s, flat):.2e} local {local(tree_s, flat):.2e}"
                      f"  auto==host-rule: {bool(np.array_equal(tree, tree_s))}", flush=True)
