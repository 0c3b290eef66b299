"""ctypes front-end of the fp64 render oracle (oracle/render_oracle.c).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see render_oracle.c header and DESIGN.md): Mitsuba 3 is not available; the oracle restates the
deterministic limit of the scene built by /root/reference/utils/mitsuba3_utils.py:92-430 and is pinned by
known-answer tests only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_SO = HERE / "_build" / "librender_oracle.so"
_lib = None

PARAM_DEFAULTS = {"metallic": 0.0, "base_R": 0.0, "base_G": 0.0, "base_B": 0.0, "roughness": 0.0, "specular": 1.0}
_NAME_TO_SLOT = {"metallic.value": 0, "base_color.value.R": 1, "base_color.value.G": 2, "base_color.value.B": 3,
                 "roughness.value": 4, "specular": 5}


def build(force: bool = False) -> Path:
    """gcc -O2 -fopenmp the C restatement into oracle/_build/ (git-ignored, travels to the GPU box)."""
    src = HERE / "render_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", str(_SO), str(src), "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.drm_oracle_render_records.argtypes = [dp, dp, ctypes.c_long, dp, dp, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_int), dp]
        _lib.drm_oracle_render_records.restype = ctypes.c_int
        _lib.drm_oracle_render_cells.argtypes = [dp, dp, ctypes.c_long, dp, dp, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_int), ctypes.c_int, dp]
        _lib.drm_oracle_render_cells.restype = ctypes.c_int
        _lib.drm_oracle_num_threads.restype = ctypes.c_int
        for fn, n in (('drm_oracle_ggx_d', 2), ('drm_oracle_smith_g1', 2), ('drm_oracle_fresnel_dielectric', 2),
                      ('drm_oracle_eta_from_specular', 1)):
            getattr(_lib, fn).restype = ctypes.c_double
            getattr(_lib, fn).argtypes = [ctypes.c_double] * n
    return _lib


def num_threads() -> int:
    return int(_load().drm_oracle_num_threads())


def set_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs ask for all host cores explicitly."""
    _load().drm_oracle_set_threads(int(n))


def z_from_named(z, names) -> np.ndarray:
    """Scene defaults (mitsuba3_utils.py:348-361) overridden by the named entries, clipped to [0,1] (:237-242)."""
    out = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 1.0])
    z = np.asarray(z, dtype=np.float64).reshape(-1)
    for value, name in zip(z, names):
        if name not in _NAME_TO_SLOT:
            raise NotImplementedError(f"BRDF parameter {name!r} is not modelled")
        out[_NAME_TO_SLOT[name]] = min(max(float(value), 0.0), 1.0)
    return out


def env_records(env: np.ndarray):
    """Texel-centre point masses of a lat-long map [He,We,3]: unit directions and radiance * solid angle (fp64).

    Convention of utils/transform.py:207-209,230-233: row 0 zenith (+Y), u = atan2(x,-z)/2pi, v = acos(y)/pi.
    """
    env = np.asarray(env, dtype=np.float64)
    He, We, _ = env.shape
    t = (np.arange(He) + 0.5) * (np.pi / He)
    p = (np.arange(We) + 0.5) * (2 * np.pi / We)
    st, ct = np.sin(t)[:, None], np.cos(t)[:, None]
    d = np.stack([st * np.sin(p)[None], np.broadcast_to(ct, (He, We)), -st * np.cos(p)[None]], -1)
    domega = (2 * np.pi / We) * (np.pi / He) * st
    E = env * domega[..., None]
    return np.ascontiguousarray(d.reshape(-1, 3)), np.ascontiguousarray(E.reshape(-1, 3))


def gauss_legendre(S: int):
    x, w = np.polynomial.legendre.leggauss(int(S))
    return np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(w / w.sum(), dtype=np.float64)


def default_alpha_min(He: int) -> float:
    """Smallest GGX alpha the texel-centre quadrature of an He-row map resolves: max(1e-3, 1.25 * pi / He)."""
    return max(1e-3, 1.25 * np.pi / He)


def render_records(dirs, E, z6, view, res, S=1, flip=False, alpha_min=1e-3, terms=3, window=None) -> np.ndarray:
    lib = _load()
    dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    z6 = np.ascontiguousarray(z6, dtype=np.float64)
    view = np.ascontiguousarray(view, dtype=np.float64)
    gx, gw = gauss_legendre(S)
    out = np.zeros((res, res, 3), np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    win = (ctypes.c_int * 4)(*[int(w) for w in window]) if window is not None else None
    rc = lib.drm_oracle_render_records(dirs.ctypes.data_as(dp), E.ctypes.data_as(dp), dirs.shape[0],
                                       z6.ctypes.data_as(dp), view.ctypes.data_as(dp), int(bool(flip)), int(res),
                                       int(S), gx.ctypes.data_as(dp), gw.ctypes.data_as(dp), float(alpha_min),
                                       int(terms), win, out.ctypes.data_as(dp))
    if rc != 0:
        raise MemoryError("render oracle allocation failed")
    return out


def render_oracle(env, z, view, res, *, names=None, S=1, flip=False, alpha_min=None, terms=3, window=None) -> np.ndarray:
    """Canonical render of one refmap: [res,res,3] fp64.  ``z`` is either the 6-vector in the shipped order
    [metallic,R,G,B,roughness,specular] (configs/drmnet/train_drmnet.yaml:26) or named by ``names``.  ``window`` =
    (i0, i1, j0, j1) restricts the evaluation to a block of cells (the rest of the output is 0)."""
    env = np.asarray(env)
    if names is None:
        names = list(_NAME_TO_SLOT)
    z6 = z_from_named(z, names)
    if alpha_min is None:
        alpha_min = default_alpha_min(env.shape[0])
    dirs, E = env_records(env)
    return render_records(dirs, E, z6, view, res, S=S, flip=flip, alpha_min=alpha_min, terms=terms, window=window)


def strided_cells(res: int, stride: int = 8) -> np.ndarray:
    """Every ``stride``-th row / column of the refmap plus the last one ([ncells,2] int32, row-major): the subset on
    which whole-image parity at the headline size is checked; it contains rows / columns 0 and res-1 (the limb)."""
    idx = sorted(set(range(0, res, stride)) | {res - 1})
    return np.array([(i, j) for i in idx for j in idx], dtype=np.int32)


def render_oracle_cells(env, z, view, res, cells, *, names=None, S=1, flip=False, alpha_min=None, terms=3) -> np.ndarray:
    """Canonical render of the listed cells only: [ncells,3] fp64 (cells = [ncells,2] rows / columns)."""
    lib = _load()
    env = np.asarray(env)
    if names is None:
        names = list(_NAME_TO_SLOT)
    z6 = np.ascontiguousarray(z_from_named(z, names))
    if alpha_min is None:
        alpha_min = default_alpha_min(env.shape[0])
    dirs, E = env_records(env)
    view = np.ascontiguousarray(view, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
    gx, gw = gauss_legendre(S)
    out = np.zeros((cells.shape[0], 3), np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib.drm_oracle_render_cells(dirs.ctypes.data_as(dp), E.ctypes.data_as(dp), dirs.shape[0],
                                     z6.ctypes.data_as(dp), view.ctypes.data_as(dp), int(bool(flip)), int(res), int(S),
                                     gx.ctypes.data_as(dp), gw.ctypes.data_as(dp), float(alpha_min), int(terms),
                                     cells.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(cells.shape[0]),
                                     out.ctypes.data_as(dp))
    if rc != 0:
        raise MemoryError("render oracle allocation failed")
    return out


def bsdf_blocks():
    """(D(n.h, alpha), G1(c, alpha), F_dielectric(cos, eta), eta(specular)) of the C restatement as Python callables."""
    lib = _load()
    return (lib.drm_oracle_ggx_d, lib.drm_oracle_smith_g1, lib.drm_oracle_fresnel_dielectric,
            lib.drm_oracle_eta_from_specular)


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


if __name__ == "__main__":
    print(build(force=True), "threads:", num_threads())
