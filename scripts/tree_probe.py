"""GPU probe of the hierarchical render: parity against the committed fp64 goldens (strided cells), against the
single-level GPU evaluation (whole image) and timing per footprint class.  Development tool (run under gpurun)."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200 import _lib  # noqa: E402
from drmnet_b200.renderer import render_batch  # noqa: E402
from drmnet_b200.synth import synthetic_envmap  # noqa: E402


def golden_check(path, max_cases=None, options=None, verbose=True):
    g = np.load(path)
    He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
    meta, vals, cells = g["meta"], g["values"], g["cells"]
    n = len(meta) if max_cases is None else min(max_cases, len(meta))
    envs = {}
    worst = 0.0
    rows = []
    for i in range(n):
        seed, zi, vi, S, nc = [int(x) for x in meta[i]]
        if seed not in envs:
            envs = {seed: torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()}
        z = torch.tensor(g["z"][i], dtype=torch.float32)[None]
        v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
        out = render_batch(envs[seed][None], z, v, res=res, footprint_S=S, alpha_min=float(g["alpha_min"]),
                           channel_first=False, options=options, check_status=True)[0].double().cpu().numpy()
        cl = cells[i][:nc]
        got, ref = out[cl[:, 0], cl[:, 1]], vals[i][:nc]
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        loc = float(np.abs(got - ref).max() / np.abs(ref).max())
        rows.append((seed, zi, vi, S, err, loc))
        worst = max(worst, err)
        if verbose:
            print(f"seed {seed} z{zi} v{vi} S={S} rough {float(g['z'][i][4]):.3f}: rel-L2 {err:.2e} max/peak {loc:.2e} "
                  f"marks {getattr(render_batch, 'last_status', None)}", flush=True)
    return worst, rows


def timing(He=1000, res=128, n=8, options=None):
    We = 2 * He
    envs = torch.stack([torch.from_numpy(synthetic_envmap(He, We, seed=1000 + b)) for b in range(n)]).cuda()
    view = torch.tensor([[0.3, 0.0, 1.0]]).repeat(n, 1)
    out = {}
    for name, rough in [("S1 r=0.7", 0.7), ("S2 r=0.3", 0.3), ("S4 r=0.15", 0.15), ("S8 r=0.09", 0.09), ("S16 r=0", 0.0)]:
        z = torch.tensor([[0.5, 0.9, 0.5, 0.3, rough, 1.0]]).repeat(n, 1)
        for _ in range(2):
            render_batch(envs, z, view, res=res, footprint_S=None, options=options)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            render_batch(envs, z, view, res=res, footprint_S=None, options=options)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3 / n
        out[name] = ms
        print(f"{name}: {ms:.3f} ms per render", flush=True)
    return out


def flat_check(He=1000, res=128, options=None):
    We = 2 * He
    for seed, z, view in [(1002, [0.5, 0.9, 0.5, 0.3, 0.3, 1.0], [-0.7, 0.0, 0.7]),
                          (1001, [0.0, 0.8, 0.6, 0.4, 0.7, 0.5], [0.3, 0.0, 1.0]),
                          (1003, [0.2, 0.7, 0.7, 0.9, 0.15, 0.8], [0.4, 0.6, 0.7])]:
        env = torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
        zt, vt = torch.tensor([z]), torch.tensor([view])
        a = render_batch(env, zt, vt, res=res, footprint_S=None, options=options, check_status=True)
        b = render_batch(env, zt, vt, res=res, footprint_S=None, flat=True)
        err = float((a - b).norm() / b.norm())
        loc = float((a - b).abs().max() / b.abs().max())
        print(f"tree vs flat seed {seed} rough {z[4]}: rel-L2 {err:.2e} max/peak {loc:.2e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["golden250", "timing", "flat", "golden1000"]
    res = {}
    if "golden250" in what:
        w, _ = golden_check(ROOT / "tests/golden/render_cells_250x500.npz", max_cases=64)
        print("golden 250x500 worst", w)
    if "timing" in what:
        res["timing"] = timing()
    if "flat" in what:
        flat_check()
    if "golden1000" in what:
        w, _ = golden_check(ROOT / "tests/golden/render_cells_1000x2000.npz")
        print("golden 1000x2000 worst", w)
    print(json.dumps(res))
