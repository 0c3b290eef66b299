#!/usr/bin/env python
"""Secondary metric: img2refmaps / s, batched, against the HBM roofline.

    python scripts/bench_img2refmap.py [--batch 64] [--radius 256] [--iters 20]

Workload: B sphere images of (2*radius)^2 pixels (~206k masked pixels each at radius 256, SURVEY 8d input (b)/(c)),
res 128, threshold pi/128/2 -- what ObsNetDiffusion.get_input does once per batch element (models/obsnet.py:318-328).
Algorithmic bytes per image: 24 n + 13 res^2 (+ 8 res^2 for counts and sel_index).  Inputs (> 300 MB) exceed L2.
"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from drmnet_b200 import _lib
from drmnet_b200.img2refmap import img2refmap_batch
from drmnet_b200.synth import sphere_image_inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--radius", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-baseline", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    res, thr = 128, float(np.pi / 128 / 2)
    cols, nrms, offs = [], [], [0]
    for b in range(a.batch):
        c, n = sphere_image_inputs(a.radius, seed=100 + b % 8)
        cols.append(c); nrms.append(n); offs.append(offs[-1] + len(c))
    colors = torch.from_numpy(np.concatenate(cols)).to(dev)
    normals = torch.from_numpy(np.concatenate(nrms)).to(dev)
    offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
    total_n = colors.shape[0]
    for _ in range(3):
        out = img2refmap_batch(colors, normals, offsets, res, thr)
    torch.cuda.synchronize()
    L = _lib.lib(); l0 = L.drm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("drm_timed")
    e0.record()
    for _ in range(a.iters):
        out = img2refmap_batch(colors, normals, offsets, res, thr)
    e1.record(); torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    ms = e0.elapsed_time(e1) / a.iters
    alg = 24 * total_n + a.batch * res * res * (13 + 8)
    peak = 6545.0
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    line = {"metric": "img2refmaps/sec (res 128, ~206k px per image)", "value": a.batch / (ms / 1e3), "unit": "img2refmaps/s",
            "ms_per_step": ms, "batch": a.batch, "pixels_per_image": total_n // a.batch, "dtype": "f32 compare / u32 keys",
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_step": alg},
            "gpu_launches": int(L.drm_launch_count() - l0), "filled_cells": int(out[1].sum())}
    if a.cpu_baseline:
        sys.path.insert(0, str(ROOT))
        from oracle.img2refmap_oracle import img2refmap_oracle
        t0 = time.perf_counter()
        img2refmap_oracle(cols[0], nrms[0], res, thr)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "img2refmaps/s", "cores": 1, "kind": "port",
                                "sample": "numpy oracle, one image of the batch"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
