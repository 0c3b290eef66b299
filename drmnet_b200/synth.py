"""Synthetic inputs for tests and benchmarks.  No reference data ships with the repo.

Envmaps follow the reference's lat-long convention (utils/transform.py:207-209,230-233 in the reference):
row 0 = zenith (+Y), column c centre at azimuth 2*pi*(c+0.5)/We measured from -Z toward +X,
``d = (sin t sin p, cos t, -sin t cos p)``.  BRDF vectors and views follow
dataset/parametricrefmap.py:105,114-116 and the schedule of models/drmnet.py:481-499.
"""
from __future__ import annotations

import math

import numpy as np
import torch

Z0 = (1.0, 1.0, 1.0, 1.0, 0.0, 1.0)  # configs/drmnet/train_drmnet.yaml:27  [metallic,R,G,B,roughness,specular]
BRDF_PARAM_NAMES = ["metallic.value", "base_color.value.R", "base_color.value.G", "base_color.value.B",
                    "roughness.value", "specular"]  # train_drmnet.yaml:26


def envmap_directions(He: int, We: int, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """Texel-centre unit directions [He,We,3] of a lat-long map in the reference's convention."""
    t = (torch.arange(He, device=device, dtype=torch.float64) + 0.5) * (math.pi / He)
    p = (torch.arange(We, device=device, dtype=torch.float64) + 0.5) * (2 * math.pi / We)
    st, ct = torch.sin(t)[:, None], torch.cos(t)[:, None]
    d = torch.stack([st * torch.sin(p)[None], ct.expand(He, We), -st * torch.cos(p)[None]], -1)
    return d.to(dtype)


def synthetic_envmap(He: int, We: int, seed: int, device="cpu", as_numpy: bool | None = None):
    """HDR-like map: lognormal low-frequency sky + 1..4 compact Gaussian lobes with 10^1..10^4 peaks.

    Random draws come from a CPU generator seeded with ``seed`` so the map is the same on every device
    up to the ulp differences of exp/sin on that device.  Returns float32 [He,We,3] (numpy when
    ``device == 'cpu'`` unless ``as_numpy`` says otherwise).
    """
    g = torch.Generator().manual_seed(int(seed))
    low = torch.exp(torch.randn(3, 32, 64, generator=g) * 0.5 + torch.randn(1, 32, 64, generator=g) * 1.0)
    tint = 0.8 + 0.4 * torch.rand(3, generator=g)
    k = int(torch.randint(1, 5, (1,), generator=g).item())
    peak = 10.0 ** (1.0 + 3.0 * torch.rand(k, generator=g))
    sigma = torch.deg2rad(0.3 + 4.7 * torch.rand(k, generator=g))
    cz = 2 * torch.rand(k, generator=g) - 1
    ca = 2 * math.pi * torch.rand(k, generator=g)
    ltint = 0.7 + 0.6 * torch.rand(k, 3, generator=g)
    cs = torch.sqrt(1 - cz * cz)
    centres = torch.stack([cs * torch.sin(ca), cz, -cs * torch.cos(ca)], -1)

    low = low.to(device)
    # bilinear up-sampling, clamped in elevation and circular in azimuth (no seam at the map edge)
    fy = ((torch.arange(He, device=device) + 0.5) * (32 / He) - 0.5).clamp(0, 31)
    fx = (torch.arange(We, device=device) + 0.5) * (64 / We) - 0.5
    y0 = fy.floor().long().clamp(0, 30)
    wy = (fy - y0).float()[None, :, None]
    x0f = fx.floor()
    wx = (fx - x0f).float()[None, None, :]
    x0 = x0f.long() % 64
    x1 = (x0 + 1) % 64
    top = low[:, y0][:, :, x0] * (1 - wx) + low[:, y0][:, :, x1] * wx
    bot = low[:, y0 + 1][:, :, x0] * (1 - wx) + low[:, y0 + 1][:, :, x1] * wx
    sky = top * (1 - wy) + bot * wy
    env = sky.permute(1, 2, 0) * tint.to(device)
    d = envmap_directions(He, We, device=device)
    for i in range(k):
        chord = torch.linalg.norm(d - centres[i].to(device), dim=-1).clamp(max=2.0)
        ang = 2 * torch.asin(chord / 2)
        lobe = peak[i].item() * torch.exp(-0.5 * (ang / sigma[i].item()) ** 2)
        env = env + lobe[..., None] * ltint[i].to(device)
    env = env.to(torch.float32).contiguous()
    if as_numpy is None:
        as_numpy = str(device) == "cpu"
    return env.cpu().numpy() if as_numpy else env


def sample_brdf(seed: int, zdim: int = 6) -> torch.Tensor:
    """zK ~ U[0,1]^zdim exactly as dataset/parametricrefmap.py:105 (torch CPU generator)."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand((zdim,), generator=g)


def sample_view(seed: int) -> torch.Tensor:
    """One of the 64 equatorial azimuths of dataset/parametricrefmap.py:114-116: (sin p, 0, cos p)."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    phi = (torch.rand((), generator=g) * 64).int() / 64 * math.pi * 2 - math.pi
    return torch.stack([torch.sin(phi), torch.zeros(()), torch.cos(phi)]).float()


def schedule_point(zK: torch.Tensor, normalized_k: float, gamma: float = 0.95, epsilon: float = 0.01,
                   z0=Z0):
    """(K, k, zk, zkm1) following models/drmnet.py:481-499 (exponent in float64)."""
    z0 = torch.tensor(z0, dtype=zK.dtype)
    delta = zK - z0
    dist = torch.linalg.norm(delta)
    K = max(int(math.log(epsilon / float(dist)) / math.log(gamma)) + 2, 1)
    k = int(normalized_k * K)
    rk = K - k - 1
    zk = z0 + float(math.exp(rk * math.log(gamma))) * delta
    zkm1 = z0 + float(math.exp((rk + 1) * math.log(gamma))) * delta
    return K, k, zk, zkm1


def sphere_image_inputs(radius: int, seed: int, noise: float = 0.05):
    """Masked (colour, normal) pixels of a sphere seen orthographically (SURVEY 8d img2refmap inputs (b),(c)).

    Normals follow gen_sphere_normals_realcentering (utils/transform.py:147-167 in the reference):
    x right, y up, z toward the viewer, then perturbed and renormalised.  Colours are a smooth HDR-like
    function of the normal plus noise.  Returns float32 numpy arrays [n,3], [n,3].
    """
    rng = np.random.default_rng(seed)
    ax = np.linspace(-radius + 0.5, radius - 0.5, 2 * radius)
    x, y = np.meshgrid(ax, -ax)
    zsq = radius ** 2 - (x ** 2 + y ** 2)
    m = zsq >= 0
    nrm = np.stack([x[m], y[m], np.sqrt(zsq[m])], -1) / radius
    nrm = nrm + noise * rng.normal(size=nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    col = np.exp(1.5 * nrm[:, :1] + 0.7 * nrm[:, 1:2]) * np.array([1.0, 0.9, 0.8]) \
        + 40.0 * np.exp(-((nrm[:, :1] - 0.3) ** 2 + (nrm[:, 1:2] - 0.5) ** 2) / 0.002)
    col = col * (1 + 0.01 * rng.normal(size=col.shape))
    return col.astype(np.float32), nrm.astype(np.float32)


def sphere_normals(radius):
    """gen_sphere_normals_realcentering (utils/transform.py:147-167) restated: x right, y up, z toward the viewer."""
    ax = np.linspace(-radius + 0.5, radius - 0.5, 2 * radius)
    x, y = np.meshgrid(ax, -ax)
    zsq = radius ** 2 - (x ** 2 + y ** 2)
    n = np.zeros((2 * radius, 2 * radius, 3), np.float32)
    n[..., 0], n[..., 1] = x, y
    n[zsq >= 0, 2] = np.sqrt(zsq[zsq >= 0])
    n /= np.sqrt((n ** 2).sum(-1, keepdims=True))
    n[zsq < 0] = 0
    ii, jj = np.ogrid[0:2 * radius, 0:2 * radius]
    mask = ((ii + 0.5 - radius) ** 2 + (jj + 0.5 - radius) ** 2) <= radius * radius
    return n * mask[..., None], mask
