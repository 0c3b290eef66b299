"""Driver of scripts/tree_proto.c (CPU replica of the hierarchical render): accuracy against the fp64 oracle on a
strided subset of cells and evaluated-pair counts, as a function of the acceptance constants.  Development tool."""
from __future__ import annotations

import ctypes
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200.renderer import auto_footprint, default_alpha_min  # noqa: E402
from drmnet_b200.synth import synthetic_envmap  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402

SO = ROOT / "scripts" / "_tree_proto.so"


def build():
    src = ROOT / "scripts" / "tree_proto.c"
    if not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", str(SO), str(src), "-lm"])
    L = ctypes.CDLL(str(SO))
    dp = ctypes.POINTER(ctypes.c_double)
    L.tree_render.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                              ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                              ctypes.c_int, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_long)]
    return L


def gl_tables():
    t = np.zeros((5, 32))
    for p in range(5):
        x, w = ro.gauss_legendre(1 << p)
        t[p, :len(x)] = x
        t[p, 16:16 + len(w)] = w
    return np.ascontiguousarray(t)


def tree_render(env, z6, view, res, S, *, alpha_min=None, terms=3, kappa=0.08, kappa_d=0.1, hz=0.03, level_scale=0.3,
                pixcov=0, full2=0, chan=0, rcap=10.0, hand=0.5, flip=False, bh=4, bw=8):
    L = build()
    env = np.ascontiguousarray(env, np.float32)
    He, We, _ = env.shape
    if alpha_min is None:
        alpha_min = default_alpha_min(He)
    z6 = np.ascontiguousarray(z6, np.float64)
    view = np.ascontiguousarray(view, np.float64)
    gl = gl_tables()
    out = np.zeros((res, res, 3))
    stats = (ctypes.c_long * 4)()
    dp = ctypes.POINTER(ctypes.c_double)
    pk = int(np.log2(S))
    L.tree_render(env.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), He, We, z6.ctypes.data_as(dp),
                  view.ctypes.data_as(dp), int(flip), res, pk, gl.ctypes.data_as(dp), alpha_min, terms, kappa, kappa_d,
                  hz, level_scale, pixcov, full2, chan, rcap, hand, bh, bw, out.ctypes.data_as(dp), stats)
    return out, dict(visits=stats[0], pairs0=stats[1], pairs1=stats[2], handed=stats[3])


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--He", type=int, default=250)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--z", type=float, nargs=6, default=[0.5, 0.9, 0.5, 0.3, 0.3, 1.0])
    ap.add_argument("--view", type=float, nargs=3, default=[0.3, 0.0, 1.0])
    ap.add_argument("--terms", type=int, default=1)
    ap.add_argument("--kappa", type=float, nargs="+", default=[0.08])
    ap.add_argument("--kappa_d", type=float, default=0.1)
    ap.add_argument("--hz", type=float, default=0.03)
    ap.add_argument("--level_scale", type=float, default=0.3)
    ap.add_argument("--pixcov", type=int, default=0)
    ap.add_argument("--full2", type=int, default=0)
    ap.add_argument("--chan", type=int, default=0)
    ap.add_argument("--rcap", type=float, default=10.0)
    ap.add_argument("--hand", type=float, default=0.5)
    ap.add_argument("--stride", type=int, default=8)
    ap.add_argument("--S", type=int, default=0)
    ap.add_argument("--bh", type=int, default=4)
    ap.add_argument("--bw", type=int, default=8)
    a = ap.parse_args()
    env = synthetic_envmap(a.He, 2 * a.He, seed=a.seed)
    amin = default_alpha_min(a.He)
    S = a.S or auto_footprint(a.z[4], a.res, amin)
    cells = ro.strided_cells(a.res, a.stride)
    t = time.time()
    ref = ro.render_oracle_cells(env, a.z, a.view, a.res, cells, S=S, alpha_min=amin, terms=a.terms)
    print(f"oracle S={S} alpha={max(a.z[4]**2, amin):.4f} {time.time() - t:.1f}s")
    for kappa in a.kappa:
        t = time.time()
        out, st = tree_render(env, a.z, a.view, a.res, S, terms=a.terms, kappa=kappa, kappa_d=a.kappa_d, hz=a.hz,
                              level_scale=a.level_scale, pixcov=a.pixcov, full2=a.full2, chan=a.chan, rcap=a.rcap, hand=a.hand, bh=a.bh, bw=a.bw)
        got = out[cells[:, 0], cells[:, 1]]
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        loc = np.abs(got - ref).max() / np.abs(ref).max()
        npx = a.res * a.res
        print(f"kappa {kappa:.3f}: rel-L2 {err:.2e} max/peak {loc:.2e} pairs/cell lvl0 {st['pairs0'] / npx:.0f} "
              f"lvl>0 {st['pairs1'] / npx:.0f} visits/cell {st['visits'] / npx:.1f} handed {st['handed']} "
              f"({time.time() - t:.1f}s)")
