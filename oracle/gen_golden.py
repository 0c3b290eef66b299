"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE's own Python code.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
The GPU box never runs this; it only reads the committed .npz files.

What is produced
* img2refmap_sample.npz   data/sample through the preprocessing of scripts/estimate.py:128-137,43-50
                          and refmap_mask_make (utils/img2refmap.py:6-37) -- BASELINE config[0].
* img2refmap_synth.npz    seeded synthetic cases covering thresholds != half a cell (multi-membership
                          and dropped pixels), min_points, NaN colours, exact ties, non-unit normals.
* mirmap_ref.npz          (also: refmap2refimg_torch on a random refmap, utils/transform.py:170-198)
                          the reference's only torch renderer, envmap2mirmap (utils/transform.py:201-242),
                          on a seeded synthetic envmap: the mirror-limit known answer, plus
                          mirmap2envmap (utils/transform.py:106-144) for the round trip (K3).
"""
from __future__ import annotations

import os
import sys
import warnings
from pathlib import Path

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
REF = Path("/root/reference")
sys.path.insert(0, str(REF))
warnings.filterwarnings("ignore")

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from utils.img2refmap import refmap_mask_make  # noqa: E402  (the reference)
from utils.transform import envmap2mirmap, mirmap2envmap, refmap2refimg_torch, xyz2thetaphi  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def sample_inputs():
    """scripts/estimate.py:128-137 (loading, mask) and :43-50 (5-px erosion), on CPU."""
    img = cv2.cvtColor(cv2.imread(str(REF / "data/sample/image.exr"), -1)[..., :3], cv2.COLOR_BGR2RGB)
    normal = np.load(REF / "data/sample/normal.npy")
    png = cv2.imread(str(REF / "data/sample/mask.png"), -1)
    img, normal = torch.from_numpy(img), torch.from_numpy(normal)
    input_mask = torch.from_numpy(png)
    if input_mask.ndim == 3:
        input_mask = input_mask[:, :, 0]
    mask = torch.logical_and(input_mask, torch.linalg.norm(normal, dim=-1) > 0.5)
    k = 5
    inv_mask = ~mask
    kernel = torch.stack(torch.meshgrid(*torch.arange(k).expand(2, -1), indexing="ij")) + 0.5
    kernel = (torch.linalg.norm(kernel - k / 2, axis=0) <= k / 2)[None, None].float()
    inv_mask = torch.nn.functional.conv2d(inv_mask[None, None].float(), kernel, padding="same").bool()[0, 0]
    mask = torch.logical_and(mask, ~inv_mask)
    return img[mask].contiguous(), normal[mask].contiguous()


def run_ref(colors, normals, res, thr, min_points=0):
    c, n = torch.from_numpy(colors), torch.from_numpy(normals)
    refmap, refmask = refmap_mask_make(c, n, res, thr, min_points=min_points)
    tp = xyz2thetaphi(n, normal=[0, 1, 0], tangent=[-1, 0, 0])
    return refmap.numpy(), refmask.numpy(), tp.numpy()


def synth_cases():
    rng = np.random.default_rng(20240607)
    cases = {}

    def sphere(n, noise):
        v = rng.normal(size=(n, 3))
        v[:, 2] = np.abs(v[:, 2])  # mostly front facing
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = v + noise * rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = v.astype(np.float32)
        v[:, 1] = np.clip(v[:, 1], -1.0, 1.0)  # keep acos finite; case G covers the NaN-angle path
        return v

    # A: half-cell threshold (what every caller uses), res 32
    nA = sphere(6000, 0.05)
    cA = rng.lognormal(0, 1, size=(6000, 3)).astype(np.float32)
    cases["A_half_cell_res32"] = (cA, nA, 32, np.pi / 32 / 2, 0)
    # B: threshold hard-wired to pi/128/2 at res 64 -> windows smaller than cells, pixels dropped
    cases["B_small_window_res64"] = (cA, nA, 64, np.pi / 128 / 2, 0)
    # C: threshold pi/128/2 at res 256 on few pixels -> overlapping windows, multi-membership
    nC = sphere(1500, 0.0)
    cC = rng.lognormal(0, 1, size=(1500, 3)).astype(np.float32)
    cases["C_overlap_res256"] = (cC, nC, 256, np.pi / 128 / 2, 0)
    # D: min_points
    cases["D_min_points_res32"] = (cA, nA, 32, np.pi / 32 / 2, 5)
    # E: NaN colours and exact ties (quantised colours), non-unit normals
    nE = sphere(4000, 0.02) * rng.uniform(0.6, 1.0, size=(4000, 1)).astype(np.float32)
    cE = np.round(rng.uniform(0, 4, size=(4000, 3))).astype(np.float32)
    cE[rng.integers(0, 4000, 60), rng.integers(0, 3, 60)] = np.nan
    cases["E_nan_ties_res16"] = (cE, nE.astype(np.float32), 16, np.pi / 16 / 2, 0)
    # G: NaN angles (|n_y| > 1): torch.amax propagates NaN, 'NaN > thr' is False -> member of every cell
    nG = sphere(800, 0.05)
    nG[[3, 77, 500], 1] = np.float32(1.0000002)
    cG = rng.lognormal(0, 1, size=(800, 3)).astype(np.float32)
    cases["G_nan_angles_res16"] = (cG, nG, 16, np.pi / 16 / 2, 0)
    # F: wide window (2.3 cells)
    cases["F_wide_window_res24"] = (cC, nC, 24, 2.3 * np.pi / 24, 3)
    return cases


def synthetic_envmap(He, We, seed):
    """Same generator as drmnet_b200.synth.synthetic_envmap, imported from the repo."""
    from drmnet_b200.synth import synthetic_envmap as gen
    return gen(He, We, seed)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    colors, normals = sample_inputs()
    refmap, refmask = refmap_mask_make(colors, normals, 128, np.pi / 128 / 2)
    tp = xyz2thetaphi(normals, normal=[0, 1, 0], tangent=[-1, 0, 0])
    print("sample: n =", colors.shape[0], "filled =", int(refmask.sum()), "sum =", float(refmap.double().sum()))
    np.savez_compressed(OUT / "img2refmap_sample.npz", colors=colors.numpy(), normals=normals.numpy(),
                        thetaphi_torch_cpu=tp.numpy(), refmap=refmap.numpy(), refmask=refmask.numpy(),
                        res=128, thr=np.pi / 128 / 2)

    blob = {}
    for name, (c, n, res, thr, mp) in synth_cases().items():
        rm, mk, tpn = run_ref(c, n, res, thr, mp)
        print(name, "filled", int(mk.sum()))
        blob.update({f"{name}/colors": c, f"{name}/normals": n, f"{name}/res": res, f"{name}/thr": thr,
                     f"{name}/min_points": mp, f"{name}/refmap": rm, f"{name}/refmask": mk,
                     f"{name}/thetaphi_torch_cpu": tpn})
    np.savez_compressed(OUT / "img2refmap_synth.npz", **blob)

    # mirror-limit known answers from the reference's torch renderer
    env = synthetic_envmap(128, 256, seed=1003)  # [He,We,3] f32
    env_t = torch.from_numpy(env).permute(2, 0, 1)[None]
    blob = {"env": env}
    for tag, view in {"v001": [0.0, 0.0, 1.0], "v100": [1.0, 0.0, 0.0],
                      "vdiag": [float(np.sin(0.7)), 0.0, float(np.cos(0.7))]}.items():
        mir = envmap2mirmap(env_t, (32, 32), view_from=view)[0].permute(1, 2, 0).numpy()
        blob[f"mirmap_{tag}"] = mir
        blob[f"view_{tag}"] = np.asarray(view, np.float32)
    mir = torch.from_numpy(blob["mirmap_v001"]).permute(2, 0, 1)[None]
    blob["envmap_from_mirmap_v001"] = mirmap2envmap(mir, (32, 64))[0].permute(1, 2, 0).numpy()
    # sphere image shaded by a refmap lookup: the reference's refmap2refimg_torch (utils/transform.py:170-198)
    g = torch.Generator().manual_seed(3)
    refmap = torch.exp(torch.randn(1, 3, 32, 32, generator=g))
    img, msk = refmap2refimg_torch(refmap, radius=24, return_mask=True)
    blob["refimg_refmap"] = refmap[0].numpy()
    blob["refimg_image"] = img[0].numpy()
    blob["refimg_mask"] = msk.numpy()
    np.savez_compressed(OUT / "mirmap_ref.npz", **blob)
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
