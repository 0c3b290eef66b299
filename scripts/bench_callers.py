#!/usr/bin/env python
"""Caller-level timings for BASELINE configs[2] and [3] (the render side only; the U-Nets are stock PyTorch, out of scope).

  config[2]  training-data synthesis: batch 20, G = 3 BRDF vectors per envmap (LrK, Lrk, Lrkm1 as DRMNet.get_input,
             models/drmnet.py:523-569) -> synthesize_refmaps = one render_batch + fused normalise/log transform
  config[3]  reverse-process step re-render: batch 32, r0 -> 128x256 envmap (r0toenvmap, models/drmnet.py:931-941) ->
             render at the current z (DRMNet.reconstruct, :943-952)
"""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.callers import r0toenvmap, refmap_postprocess, synthesize_refmaps
from drmnet_b200.renderer import B200RefMapRenderer, render_batch
from drmnet_b200.synth import BRDF_PARAM_NAMES, Z0, sample_brdf, sample_view, schedule_point, synthetic_envmap

dev = "cuda:0"


def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    # ---- config[2]: training-data synthesis --------------------------------------------------------------------
    B = 20
    envs = torch.stack([synthetic_envmap(1000, 2000, 2000 + b, device=dev, as_numpy=False) for b in range(B)])
    zK = torch.stack([sample_brdf(2000 + b) for b in range(B)])
    sched = [schedule_point(zK[b], float(torch.rand((), generator=torch.Generator().manual_seed(b)))) for b in range(B)]
    stacked_z = torch.stack([zK, torch.stack([s[2] for s in sched]).float(), torch.stack([s[3] for s in sched]).float()])
    views = torch.stack([sample_view(2000 + b) for b in range(B)])
    r = B200RefMapRenderer(refmap_res=128, spp=256, denoise="simple", brdf_param_names=BRDF_PARAM_NAMES)
    ms = timed(lambda: synthesize_refmaps(r, stacked_z, envs, views), 3)
    out["config2_training_synthesis"] = {"batch": B, "renders_per_step": 3 * B, "ms_per_step": ms,
                                         "refmaps_per_s": 3 * B / (ms / 1e3)}
    # ---- config[3]: per-step re-render during sampling ------------------------------------------------------
    B = 32
    basis = render_batch(torch.ones(1, 128, 256, 3, device=dev), torch.tensor([list(Z0)]), torch.tensor([[0.0, 0.0, 1.1]]),
                         res=128, footprint_S=2)[0]
    r0 = render_batch(envs[:1], torch.tensor([list(Z0)]), torch.tensor([[0.0, 0.0, 1.1]]), res=128, footprint_S=4)
    r0 = r0.expand(B, 3, 128, 128).contiguous()
    z = torch.stack([sample_brdf(3000 + b) for b in range(B)])
    view = torch.tensor([[0.0, 0.0, 1.1]]).expand(B, 3)

    def step():
        env = r0toenvmap(r0, basis.clamp_min(1e-3), (128, 256)).contiguous()
        return render_batch(env, z, view, res=128, footprint_S=None)
    ms = timed(step, 5)
    out["config3_sampling_rerender"] = {"batch": B, "envmap": "128x256 from r0toenvmap", "ms_per_step": ms,
                                        "refmaps_per_s": B / (ms / 1e3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
