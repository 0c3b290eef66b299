"""CPU oracle for the image -> refmap scatter.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (drmnet_b200/) never does.

Restates /root/reference/utils/img2refmap.py:6-37 (``refmap_mask_make``) and the part of
/root/reference/utils/transform.py:55-89 (``xyz2thetaphi``) it calls at img2refmap.py:20, as an
O(n * window) numpy program instead of the reference's O(res^2 * n) dense window test.

Semantics followed (all comparisons in float32, exactly like the torch code):

* cell centres  theta_i = fp32(i + 0.5) * fp32(pi / res)  (img2refmap.py:16-17; the python scalar
  pi/res is rounded to fp32 once, the product is one fp32 multiply),
* pixel angles  theta = acos(n . [0,1,0]), phi = atan2(n . [0,0,1], n . [-1,0,0])
  (transform.py:84-89 with normal=[0,1,0], tangent=[-1,0,0]; binormal = cross = [0,0,1]),
* membership    NOT (max(|theta_i - theta_p|, |phi_j - phi_p|) > fp32(thr))   (img2refmap.py:26-27);
  a NaN angle makes the comparison False, so such a pixel is a member of EVERY cell (torch.amax
  propagates NaN) -- restated faithfully,
* a cell with fewer than ``min_points`` members is empty (img2refmap.py:28; the count includes
  members whose colour sum is NaN),
* otherwise the member with the LOWER MEDIAN of colors.sum(-1) among non-NaN sums is selected
  (torch.nanmedian, img2refmap.py:30-31: rank (cnt_valid - 1) // 2 ascending) and its whole colour
  vector is copied (img2refmap.py:34); the channel sum is ((c0 + c1) + c2) in fp32 (verified against
  torch on data/sample by tests/test_oracle_img2refmap.py),
* ties on the sum: torch leaves the choice implementation-defined; this project defines it as the
  lowest pixel index.  The selected SUM is identical either way.

Extra outputs (not returned by the reference, derivable from it): per-cell member ``counts`` and the
selected pixel index ``sel_index`` (-1 for empty cells).  ``reduce="mean"`` is an additive mode
(fp32 sum of member colours in ascending pixel order, divided by the valid count).
"""
from __future__ import annotations

import numpy as np

__all__ = ["thetaphi_from_normals", "cell_centres", "img2refmap_oracle", "img2refmap_batch_oracle"]


def thetaphi_from_normals(normals: np.ndarray) -> np.ndarray:
    """[n,3] float32 -> [n,2] float32 (theta, phi); transform.py:84-89 as called at img2refmap.py:20.

    The three-term dot products with 0 / +-1 coefficients are evaluated literally so that -0.0
    inputs normalise to +0.0 exactly as the torch matmul does.
    """
    n = np.ascontiguousarray(normals, dtype=np.float32)
    zero = np.float32(0.0)
    with np.errstate(invalid="ignore"):
        ny = n[:, 0] * zero + n[:, 1] * np.float32(1.0) + n[:, 2] * zero
        nt = n[:, 0] * np.float32(-1.0) + n[:, 1] * zero + n[:, 2] * zero
        nb = n[:, 0] * zero + n[:, 1] * zero + n[:, 2] * np.float32(1.0)
        # correctly-rounded fp32 angles (double evaluation, one rounding): libm/SLEEF/CUDA acosf and
        # atan2f each sit within a few ulp of this; see tests for the measured distances
        theta = np.arccos(ny.astype(np.float64)).astype(np.float32)
        phi = np.arctan2(nb.astype(np.float64), nt.astype(np.float64)).astype(np.float32)
    return np.stack([theta, phi], -1)


def cell_centres(res: int) -> np.ndarray:
    """fp32 centres (i + 0.5) * (pi / res), img2refmap.py:16-17."""
    return (np.arange(res).astype(np.float32) + np.float32(0.5)) * np.float32(np.pi / res)


def _members(thetaphi: np.ndarray, res: int, thr: np.float32):
    """Return (cell, pixel) membership pairs, pixel-major order, plus the NaN-angle pixel list."""
    centres = cell_centres(res)
    step = np.float64(np.pi / res)
    th, ph = thetaphi[:, 0], thetaphi[:, 1]
    nan_px = np.nonzero(np.isnan(th) | np.isnan(ph))[0]
    ok = ~(np.isnan(th) | np.isnan(ph))
    idx = np.nonzero(ok)[0]
    R = int(np.ceil(float(thr) / step)) + 1
    # candidate generator: floor bin +- R; the fp32 predicate below is the only thing that decides
    i0 = np.floor(th[idx].astype(np.float64) / step).astype(np.int64)
    j0 = np.floor(ph[idx].astype(np.float64) / step).astype(np.int64)
    cells, pixels = [], []
    offs = np.arange(-R, R + 1)
    for di in offs:
        ii = i0 + di
        vi = (ii >= 0) & (ii < res)
        dth = np.abs(centres[np.clip(ii, 0, res - 1)] - th[idx])  # fp32 - fp32
        in_i = vi & ~(dth > thr)
        if not in_i.any():
            continue
        for dj in offs:
            jj = j0 + dj
            vj = (jj >= 0) & (jj < res)
            dph = np.abs(centres[np.clip(jj, 0, res - 1)] - ph[idx])
            m = in_i & vj & ~(dph > thr)
            if m.any():
                cells.append(ii[m] * res + jj[m])
                pixels.append(idx[m])
    if cells:
        cells = np.concatenate(cells)
        pixels = np.concatenate(pixels)
    else:
        cells = np.zeros(0, np.int64)
        pixels = np.zeros(0, np.int64)
    return cells, pixels, nan_px


def img2refmap_oracle(colors, normals, res, angle_threshold, min_points=0, *, thetaphi=None, reduce="median"):
    """Oracle for ``refmap_mask_make`` (img2refmap.py:6-37).

    colors [n,C] float32, normals [n,3] float32 (or ``thetaphi`` [n,2] float32 to bypass acos/atan2).
    Returns (refmap [res,res,C] f32, refmask [res,res] bool, counts [res,res] i32, sel_index [res,res] i32).
    """
    colors = np.ascontiguousarray(colors, dtype=np.float32)
    n, C = colors.shape
    if n == 0:
        # torch.nanmedian raises IndexError on an empty reduction dim (img2refmap.py:31)
        raise IndexError("median(): Expected reduction dim 1 to have non-zero size.")
    if angle_threshold is None:
        raise TypeError("'>' not supported between instances of 'Tensor' and 'NoneType'")
    thr = np.float32(angle_threshold)
    if thetaphi is None:
        thetaphi = thetaphi_from_normals(normals)
    thetaphi = np.ascontiguousarray(thetaphi, dtype=np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        s = colors[:, 0].copy()
        for c in range(1, C):
            s = (s + colors[:, c]).astype(np.float32)  # ((c0 + c1) + c2) in fp32
    cells, pixels, nan_px = _members(thetaphi, res, thr)
    if nan_px.size:  # NaN-angle pixels are members of every cell
        cells = np.concatenate([cells, np.repeat(np.arange(res * res), nan_px.size)])
        pixels = np.concatenate([pixels, np.tile(nan_px, res * res)])
    counts = np.bincount(cells, minlength=res * res).astype(np.int32)

    refmap = np.zeros((res * res, C), np.float32)
    refmask = np.zeros(res * res, bool)
    sel = np.full(res * res, -1, np.int32)

    valid = ~np.isnan(s[pixels])
    cv, pv = cells[valid], pixels[valid]
    # total order inside a cell: (sum value, pixel index); -0.0 == +0.0 compare equal like torch
    order = np.lexsort((pv, s[pv], cv))
    cv, pv = cv[order], pv[order]
    starts = np.searchsorted(cv, np.arange(res * res), side="left")
    ends = np.searchsorted(cv, np.arange(res * res), side="right")
    nvalid = ends - starts
    filled = (nvalid > 0) & (counts >= min_points)
    cell_ids = np.nonzero(filled)[0]
    if reduce == "median":
        pick = pv[starts[cell_ids] + (nvalid[cell_ids] - 1) // 2]
        refmap[cell_ids] = colors[pick]
        sel[cell_ids] = pick
    elif reduce == "mean":
        for c in cell_ids:
            px = np.sort(pv[starts[c]:ends[c]])
            acc = np.zeros(C, np.float32)
            for p in px:
                acc = (acc + colors[p]).astype(np.float32)
            refmap[c] = acc / np.float32(px.size)
    else:
        raise ValueError(reduce)
    refmask[cell_ids] = True
    return (refmap.reshape(res, res, C), refmask.reshape(res, res),
            counts.reshape(res, res), sel.reshape(res, res))


def img2refmap_batch_oracle(colors, normals, offsets, res, angle_threshold, min_points=0, *, thetaphi=None,
                            reduce="median"):
    """Segmented form: image b owns rows offsets[b]:offsets[b+1].  sel_index is image-local."""
    outs = []
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        if hi == lo:  # an empty image yields an empty refmap in the batched API (documented deviation)
            C = np.asarray(colors).shape[1]
            outs.append((np.zeros((res, res, C), np.float32), np.zeros((res, res), bool),
                         np.zeros((res, res), np.int32), np.full((res, res), -1, np.int32)))
            continue
        tp = None if thetaphi is None else thetaphi[lo:hi]
        outs.append(img2refmap_oracle(colors[lo:hi], None if normals is None else normals[lo:hi], res,
                                      angle_threshold, min_points, thetaphi=tp, reduce=reduce))
    return tuple(np.stack([o[k] for o in outs]) for k in range(4))
