"""Small invocations of every kernel of libdrmrender.so for compute-sanitizer runs (memcheck / racecheck), GPU:
the hierarchical render with every footprint (one mixed launch sequence and per-footprint calls), the single-level
validation kernel, img2refmap (median and mean), and the post-processing / warp kernels."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from drmnet_b200.callers import (mirmap2envmap, normalized_log_apply, normalized_log_rescale, normalized_log_transform,
                                  obsnet_condition, refmap_lookup, refmap_postprocess)
from drmnet_b200.img2refmap import img2refmap_batch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import Z0, sphere_image_inputs, synthetic_envmap

dev = "cuda:0"
env = synthetic_envmap(125, 250, 3, device=dev, as_numpy=False)[None]
zs = [list(Z0), [1.0, 0.9, 0.8, 0.7, 0.1, 1.0], [0.5, 0.9, 0.8, 0.7, 0.2, 1.0], [0.3, 0.8, 0.6, 0.4, 0.4, 0.7],
      [0.3, 0.8, 0.6, 0.4, 0.9, 0.7]]
view = torch.tensor([[0.3, 0.0, 1.0]])
o = render_batch(env, torch.tensor(zs), view.expand(5, 3), env_index=torch.zeros(5, dtype=torch.int32), res=20,
                 footprint_S=[16, 8, 4, 2, 1], check_status=True)
print("mixed footprints", [round(float(x.mean()), 5) for x in o], render_batch.last_status[:8])
for S, z in zip((16, 8, 4, 2, 1), zs):
    a = render_batch(env, torch.tensor([z]), view, res=20, footprint_S=S, check_status=True)
    print("S", S, float(a.mean()))
a = render_batch(env, torch.tensor([zs[3]]), view, res=20, footprint_S=None)
b = render_batch(env, torch.tensor([zs[3]]), view, res=20, footprint_S=3, flat=True)
print("auto", float(a.mean()), "flat S=3", float(b.mean()))
c, n = sphere_image_inputs(24, seed=2)
offs = torch.tensor([0, len(c) // 2, len(c)], dtype=torch.int64, device=dev)
for mode in ("median", "mean"):
    r = img2refmap_batch(torch.from_numpy(c).to(dev), torch.from_numpy(n).to(dev), offs, 16, float(np.pi / 32), reduce=mode,
                         check_status=True)
    print("img2refmap", mode, int(r[1].sum()), img2refmap_batch.last_status)
# windows wider than a cell (several cells per pixel), cells larger than the staging buffer (batched select, warp-per-cell path)
for res, thr in ((24, 0.2), (4, float(np.pi / 8))):
    r = img2refmap_batch(torch.from_numpy(c).to(dev), torch.from_numpy(n).to(dev), offs, res, thr, check_status=True)
    print("img2refmap res", res, "thr", round(thr, 3), int(r[1].sum()), img2refmap_batch.last_status)
c2, n2 = sphere_image_inputs(96, seed=4)
r = img2refmap_batch(torch.from_numpy(c2).to(dev), torch.from_numpy(n2).to(dev), torch.tensor([0, len(c2)]), 16, float(np.pi / 32),
                     check_status=True)
print("img2refmap ~110 px per cell", int(r[1].sum()), img2refmap_batch.last_status)
st = torch.rand(3, 2, 3, 16, 16, device=dev)
print("post", float(refmap_postprocess(st)[0].mean()))
print("warp", float(mirmap2envmap(st[0], (16, 32)).mean()))
print("lookup", float(refmap_lookup(st[0, :1], torch.from_numpy(n).to(dev)).mean()))
t, prm = normalized_log_transform(st[0] + 0.1, torch.ones(2, 1, 16, 16, device=dev))
print("nlog", float(t.mean()), float(normalized_log_rescale(normalized_log_apply(st[1] + 0.1, prm), prm).mean()))
print("cond", float(obsnet_condition(st[0] + 0.1, torch.rand(2, 16, 16, device=dev) > 0.5, noisy_observe=0.1,
                                     padding_mode="noise")[0].mean()))
torch.cuda.synchronize()
