#!/usr/bin/env python
"""BASELINE config[4]: resolution / envmap-size / batch sweep of the batched render (one JSON line per point).

    python scripts/bench_sweep.py [--quick]

z ~ U[0,1]^6 per render, footprint chosen per render (auto), one envmap per render.  Each point is checked for finite
output; timing = CUDA events around 2 steps after 1 warm-up.
"""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import sample_brdf, sample_view, synthetic_envmap

dev = "cuda:0"
quick = "--quick" in sys.argv
points = [(res, env, n) for res in (64, 128, 256) for env in ((250, 500), (500, 1000), (1000, 2000)) for n in (1, 8, 64)]
if quick:
    points = [(64, (250, 500), 8), (128, (500, 1000), 8), (256, (1000, 2000), 8), (128, (1000, 2000), 1), (128, (250, 500), 64)]
points.append((128, (250, 500), 512))
cache = {}
for res, (He, We), n in points:
    key = (He, We)
    nb = min(n, 16)  # distinct envmaps (renders cycle over them): keeps the sweep's memory bounded
    if key not in cache:
        cache[key] = torch.stack([synthetic_envmap(He, We, 5000 + b, device=dev, as_numpy=False) for b in range(16)])
    envs = cache[key][:nb]
    z = torch.stack([sample_brdf(5000 + i) for i in range(n)])
    v = torch.stack([sample_view(5000 + i) for i in range(n)])
    idx = torch.arange(n) % nb
    f = lambda: render_batch(envs, z, v, env_index=idx, res=res, footprint_S=None)
    out = f(); torch.cuda.synchronize()
    assert torch.isfinite(out).all() and out.shape == (n, 3, res, res)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(json.dumps({"res": res, "envmap": f"{We}x{He}", "batch": n, "ms_per_step": round(ms, 2),
                      "refmaps_per_s": round(n / (ms / 1e3), 2)}), flush=True)
