"""Accuracy (against the fp64 golden cells at 2000x1000) and time of the sharp footprints under different rim options.
usage: rim_probe.py "name=limb_sub:-1e-9,limb_hand:2" ...   (development tool, GPU)"""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200 import _lib
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap

GOLD = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("file=")]
sys.argv = [a for a in sys.argv if not a.startswith("file=")]
g = np.load(ROOT / "tests/golden" / (GOLD[0] if GOLD else "render_cells_1000x2000.npz"))
He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
meta, vals, cells = g["meta"], g["values"], g["cells"]
want_S = (1, 2, 4, 8, 16)
idx = [i for i in range(len(meta)) if int(meta[i][3]) in want_S]
variants = [("warm-up", {}), ("default", {})]
for a in sys.argv[1:]:
    name, kv = a.split("=", 1)
    variants.append((name, {k: float(v) for k, v in (x.split(":") for x in kv.split(","))}))
envs = {}
for name, kw in variants:
    o = _lib.default_render_options()
    for k, v in kw.items():
        setattr(o, k, type(getattr(o, k))(v))
    errs, locs, t_ms = {}, {}, {}
    for i in idx:
        seed, zi, vi, S, nc = [int(x) for x in meta[i]]
        if seed not in envs:
            envs[seed] = torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
        z = torch.tensor(g["z"][i], dtype=torch.float32)[None]
        v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
        kwargs = dict(res=res, footprint_S=S, alpha_min=float(g["alpha_min"]), channel_first=False, options=o)
        out = render_batch(envs[seed], z, v, **kwargs)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        try:
            out = render_batch(envs[seed], z, v, check_status=True, **kwargs)[0].double().cpu().numpy()
        except Exception as e:
            print("   case", i, "S", S, "failed:", str(e)[:100]); continue
        t_ms.setdefault(S, []).append((time.perf_counter() - t0) * 1e3)
        cl = cells[i][:nc]
        got, ref = out[cl[:, 0], cl[:, 1]], vals[i][:nc]
        errs.setdefault(S, []).append(float(np.linalg.norm(got - ref) / np.linalg.norm(ref)))
        locs.setdefault(S, []).append(float(np.abs(got - ref).max() / np.abs(ref).max()))
        if errs[S][-1] > (4e-4 if S == 16 else 1e-4) or locs[S][-1] > 1e-3:
            d = np.abs(got - ref).max(axis=1)
            top = np.argsort(-d)[:6]
            print(f"   FAIL case {i}: env {seed} z{zi} {np.round(g['z'][i], 3).tolist()} view{vi} {np.round(g['view'][i], 3).tolist()} S={S} "
                  f"rel-L2 {errs[S][-1]:.2e} loc {locs[S][-1]:.2e}; worst cells (row, col, |err|/peak, got/ref of the max channel): "
                  + "; ".join(f"({cl[t,0]},{cl[t,1]}) {d[t]/np.abs(ref).max():.1e} {got[t].max():.4g}/{ref[t].max():.4g}" for t in top))
    print(f"[{name}] {kw}")
    for S in sorted(errs):
        e, l, t = np.array(errs[S]), np.array(locs[S]), np.array(t_ms[S])
        tol = 4e-4 if S == 16 else 1e-4
        print(f"  S={S:2d} n={len(e):3d}  rel-L2 max {e.max():.2e} mean {e.mean():.2e} (fail {int((e > tol).sum())})  "
              f"worst cell/peak max {l.max():.2e} (fail {int((l > 1e-3).sum())})  time mean {t.mean():6.2f} ms")
