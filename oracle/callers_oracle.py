"""CPU oracles for the callers either side of the render.  TEST INFRASTRUCTURE ONLY.

* ``postprocess_oracle`` restates models/drmnet.py:610-620 + dataset/basedataset.py:52-53 (float64).
* ``mirmap2envmap_oracle`` restates utils/transform.py:106-144 (defaults) + torch.nn.functional.grid_sample
  (bilinear, padding_mode="border", align_corners=False) in float64; pinned by tests/golden/mirmap_ref.npz, which holds
  the reference's own output.
"""
import numpy as np


def postprocess_oracle(stacks, target=0.12, transform="log"):
    x = np.asarray(stacks, dtype=np.float64)
    G, N = x.shape[:2]
    scale = np.ones(N)
    if target:
        r = x[0]
        L = 0.212671 * r[:, 0] + 0.715160 * r[:, 1] + 0.072169 * r[:, 2]
        m = L > 0
        Lmean = np.exp((np.log(np.clip(L, 1e-5, None)) * m).sum((1, 2)) / m.sum((1, 2)))
        scale = target / Lmean
    y = x * scale[None, :, None, None, None]
    if transform == "log":
        y = np.log10(y + 0.1) + 1
    return y, scale


def mirmap2envmap_oracle(mirmap, output_shape, basis=None):
    m = np.asarray(mirmap, dtype=np.float64)
    if basis is not None:
        m = m / np.asarray(basis, dtype=np.float64)[None]
    B, C, H, W = m.shape
    OH, OW = output_shape
    theta = (np.arange(OH) + 0.5) * (np.pi / OH)
    phi = -(np.arange(OW) + 0.5) * (2 * np.pi / OW)
    th, ph = np.meshgrid(theta, phi, indexing="ij")
    # thetaphi2xyz(normal=[0,1,0], tangent=[0,0,-1]) -> binormal [-1,0,0]
    xyz = np.stack([-np.sin(th) * np.sin(ph), np.cos(th), -np.sin(th) * np.cos(ph)], -1)
    h = xyz + np.array([0.0, 0.0, 1.0])
    h = h / np.clip(np.linalg.norm(h, axis=-1, keepdims=True), 1e-12, None)
    t_h = np.arccos(np.clip(h[..., 1], -1, 1))
    p_h = np.arctan2(h[..., 0], h[..., 2])
    u, v = p_h * (2 / np.pi), t_h * (2 / np.pi) - 1
    ix = np.clip(((u + 1) * W - 1) / 2, 0, W - 1)
    iy = np.clip(((v + 1) * H - 1) / 2, 0, H - 1)
    x0, y0 = np.floor(ix).astype(int), np.floor(iy).astype(int)
    wx, wy = ix - x0, iy - y0
    x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)  # weight is 0 whenever the clamp is active
    out = (m[:, :, y0, x0] * (1 - wx) * (1 - wy) + m[:, :, y0, x1] * wx * (1 - wy)
           + m[:, :, y1, x0] * (1 - wx) * wy + m[:, :, y1, x1] * wx * wy)
    return out


def refmap_lookup_oracle(refmap, normals):
    """refmap [C,H,W], normals [n,3] -> [n,C]: utils/transform.py:170-198 (uv from xyz2thetaphi(n,[0,1,0],[-1,0,0]),
    grid_sample bilinear / border / align_corners=False), float64."""
    m = np.asarray(refmap, dtype=np.float64)
    n = np.asarray(normals, dtype=np.float64)
    C, H, W = m.shape
    th = np.arccos(np.clip(n[:, 1], -1, 1))
    ph = np.arctan2(n[:, 2], -n[:, 0] + 0.0)
    u, v = ph * (2 / np.pi) - 1, th * (2 / np.pi) - 1
    ix = np.clip(((u + 1) * W - 1) / 2, 0, W - 1)
    iy = np.clip(((v + 1) * H - 1) / 2, 0, H - 1)
    x0, y0 = np.floor(ix).astype(int), np.floor(iy).astype(int)
    wx, wy = ix - x0, iy - y0
    x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)
    out = (m[:, y0, x0] * (1 - wx) * (1 - wy) + m[:, y0, x1] * wx * (1 - wy)
           + m[:, y1, x0] * (1 - wx) * wy + m[:, y1, x1] * wx * wy)
    return out.T


def normalized_log_oracle(x, mask, lowerbound=1e-6):
    """dataset/basedataset.py:56-76 for `0p1tom1p1_normalizedLogarithmic_lowerbound<lb>` with dynamic_normalize."""
    x = np.clip(np.asarray(x, dtype=np.float64), lowerbound, None)
    m = np.asarray(mask, dtype=np.float64).reshape(x.shape[0], 1, x.shape[2], x.shape[3])
    linearmax = (x * m).max(axis=(1, 2, 3), keepdims=True)
    log10max = np.log10(linearmax)
    log10min = np.log10((x * m + (1 - m) * linearmax).min(axis=(1, 2, 3), keepdims=True))
    y = (np.log10(x) - log10min) / (log10max - log10min)
    return y * 2 - 1, log10min.reshape(-1), log10max.reshape(-1)


def normalized_log_apply_oracle(x, log10min, log10max, lowerbound=1e-6):
    """dataset/basedataset.py:68-72 with stored parameters (dynamic_normalize=False), then `0p1tom1p1`."""
    x = np.clip(np.asarray(x, dtype=np.float64), lowerbound, None)
    a = np.asarray(log10min, dtype=np.float64).reshape(-1, 1, 1, 1)
    b = np.asarray(log10max, dtype=np.float64).reshape(-1, 1, 1, 1)
    return (np.log10(x) - a) / (b - a) * 2 - 1


def normalized_log_rescale_oracle(y, log10min, log10max, clamp_before_exp=0.0):
    """BaseDataset.rescale for the same chain (dataset/basedataset.py:83-110)."""
    a = np.asarray(log10min, dtype=np.float64).reshape(-1, 1, 1, 1)
    b = np.asarray(log10max, dtype=np.float64).reshape(-1, 1, 1, 1)
    t = (np.asarray(y, dtype=np.float64) + 1) / 2 * (b - a) + a
    if clamp_before_exp:
        t = np.minimum(t, clamp_before_exp)
    return 10.0 ** t


def obsnet_condition_oracle(raw_refmap, raw_refmask, lowerbound=1e-6, noisy_observe=0.0, observe_noise=None,
                            padding_noise=None):
    """models/obsnet.py:672-691 for cond_stage_key "raw_refmap" (image_size equal to the refmap size)."""
    t, lmin, lmax = normalized_log_oracle(raw_refmap, raw_refmask, lowerbound)
    m = np.asarray(raw_refmask, dtype=np.float64).reshape(t.shape[0], 1, t.shape[2], t.shape[3])
    cond = t * m
    if noisy_observe > 0:
        cond = noisy_observe * np.asarray(observe_noise, dtype=np.float64) + cond
    if padding_noise is not None:
        cond = cond + (1 - m) * np.asarray(padding_noise, dtype=np.float64)
    return cond, lmin, lmax
