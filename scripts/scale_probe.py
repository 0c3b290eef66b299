"""Why does a LARGER level scale hurt at res 128 / S 16 on a small map?  Toggle stages at scale 3 (GPU)."""
import os, sys
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap, Z0
dev = "cuda:0"
env = synthetic_envmap(250, 500, 1004, device=dev, as_numpy=False)[None]
v = torch.tensor([[0.6442, 0.0, 0.7648]])
SW = ["DRM_RENDER_COARSE", "DRM_RENDER_NEAR", "DRM_RENDER_FAR_COARSE", "DRM_RENDER_VIEW_AVG"]
def run(res, S, levels, scale, off=()):
    for k in SW: os.environ.pop(k, None)
    for k in off: os.environ[k] = "0"
    os.environ["DRM_RENDER_LEVELS"] = "1" if levels else "0"; os.environ["DRM_RENDER_LEVEL_SCALE"] = str(scale)
    return render_batch(env, torch.tensor([list(Z0)]), v, res=res, footprint_S=S)
res, S = 128, 16
ref = run(res, S, False, 1)
for sc in (0.3, 1.0, 2.0, 3.0):
    for off in ((), ("DRM_RENDER_NEAR",), ("DRM_RENDER_VIEW_AVG",), ("DRM_RENDER_FAR_COARSE",), ("DRM_RENDER_COARSE",)):
        o = run(res, S, True, sc, off)
        err = ((o - ref).norm() / ref.norm()).item()
        e = (o - ref).abs().amax(1)[0]; i = int(e.argmax())
        print(f"scale {sc} off {off}: {err:.2e} max at {i//res},{i%res}: out {o[0,:,i//res,i%res].tolist()} ref {ref[0,:,i//res,i%res].tolist()}")
