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


def smooth_mirmap_golden():
    """Mirror-limit known answer on a SMOOTH map (three broad lobes + a floor): the reference's envmap2mirmap output for
    two views.  tests/test_oracle_brdf_pins.py renders z0 on a super-sampled copy of this map and compares at 1e-2."""
    He, We = 64, 128
    t = (np.arange(He) + 0.5) * np.pi / He
    p = (np.arange(We) + 0.5) * 2 * np.pi / We
    st, ct = np.sin(t)[:, None], np.cos(t)[:, None]
    d = np.stack([st * np.sin(p)[None], np.broadcast_to(ct, (He, We)), -st * np.cos(p)[None]], -1)
    env = np.zeros((He, We, 3))
    for k, amp, col in (((0.3, 0.5, 0.8), 3.0, (1, .8, .6)), ((-0.7, 0.2, -0.4), 2.0, (.5, .9, 1.0)),
                        ((0.1, -0.8, 0.5), 1.5, (.9, .9, .4))):
        k = np.array(k) / np.linalg.norm(k)
        env += amp * np.exp(4.0 * (d @ k - 1))[..., None] * np.array(col)
    env = (env + 0.2).astype(np.float32)
    env_t = torch.from_numpy(env).permute(2, 0, 1)[None]
    blob = {"env": env}
    for tag, view in {"v001": [0.0, 0.0, 1.0], "vdiag": [0.6, 0.0, 0.8]}.items():
        blob[f"mirmap_{tag}"] = envmap2mirmap(env_t, (32, 32), view_from=view)[0].permute(1, 2, 0).numpy()
        blob[f"view_{tag}"] = np.asarray(view, np.float32)
    np.savez_compressed(OUT / "mirmap_smooth.npz", **blob)


def callers_golden():
    """Known answers for the steps either side of the render, produced by the reference's own code:
    * dataset/basedataset.py BaseDataset.transform for `log` and for ObsNet's
      `0p1tom1p1_normalizedLogarithmic_lowerbound1e-6` with dynamic normalisation under a mask (:52-76), imported;
    * the luminance normalisation of DRMNet.get_input (models/drmnet.py:610-617): models/drmnet.py cannot be imported
      here (mitsuba, pytorch_lightning), so exactly those source lines are read from the reference file and executed."""
    import textwrap
    import types
    from dataset.basedataset import BaseDataset
    g = torch.Generator().manual_seed(11)
    x = torch.exp(1.5 * torch.randn(3, 4, 3, 16, 16, generator=g))  # [G, N, 3, res, res]
    x[0, 1, :, :4] = 0.0  # zero-luminance pixels are left out of the geometric mean
    ds = BaseDataset(size=16, transform_func="log")
    src = (REF / "models" / "drmnet.py").read_text().splitlines()[609:617]
    assert src[0].strip().startswith("if self.refmap_input_scaler is not None:") and "stacked_Lr[idx] = Lr" in src[-1], src
    ns = {"torch": torch, "self": types.SimpleNamespace(refmap_input_scaler=0.12), "stacked_Lr": [t.clone() for t in x]}
    exec(textwrap.dedent("\n".join(src)), ns)
    scaled = torch.stack(ns["stacked_Lr"])
    out = torch.stack([ds.transform(t) for t in scaled])
    blob = {"post_in": x.numpy(), "post_scale": ns["self"].normalizing_scale.numpy(), "post_out": out.numpy()}
    ds2 = BaseDataset(size=16, transform_func="0p1tom1p1_normalizedLogarithmic_lowerbound1e-6")
    y = torch.exp(2.0 * torch.randn(4, 3, 16, 16, generator=g))
    y[0, :, 0, 0] = 0.0  # below the lower bound
    mask = (torch.rand(4, 1, 16, 16, generator=g) > 0.3)
    blob["nlog_in"] = y.numpy()
    blob["nlog_mask"] = mask.numpy()
    blob["nlog_out"] = ds2.transform(y, dynamic_normalize=True, mask=mask).numpy()
    blob["nlog_min"] = ds2.Logarithmic_params[0].reshape(-1).numpy()
    blob["nlog_max"] = ds2.Logarithmic_params[1].reshape(-1).numpy()
    # the same chain with stored parameters (dynamic_normalize=False, LrK at models/obsnet.py:371) and its inverse
    z = torch.exp(2.0 * torch.randn(4, 3, 16, 16, generator=g))
    blob["nlog_fixed_in"] = z.numpy()
    blob["nlog_fixed_out"] = ds2.transform(z, dynamic_normalize=False, mask=mask).numpy()
    w = torch.rand(4, 3, 16, 16, generator=g) * 2.4 - 1.2
    blob["nlog_rescale_in"] = w.numpy()
    blob["nlog_rescale_out"] = ds2.rescale(w).numpy()
    ds3 = BaseDataset(size=16, transform_func="0p1tom1p1_normalizedLogarithmic_lowerbound1e-6", clamp_before_exp=0.5)
    ds3.Logarithmic_params = ds2.Logarithmic_params
    blob["nlog_rescale_clamped_out"] = ds3.rescale(w).numpy()
    # ObsNet.get_cond_for_predict (models/obsnet.py:663-695): models/obsnet.py cannot be imported here either, so the
    # source lines are executed with a stand-in `self`; torch.randn_like is wrapped to record the noise it drew
    osrc = (REF / "models" / "obsnet.py").read_text().splitlines()[662:695]
    assert osrc[0].strip() == "if self.model.conditioning_key is not None:" and "raise NotImplementedError()" in osrc[-1], osrc
    drawn = []

    class _T:  # torch with a recording randn_like
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def randn_like(t):
            n = torch.randn(t.shape, generator=g)
            drawn.append(n)
            return n

    for tag, sigma, padding in (("plain", 0.0, "zeros"), ("noisy", 0.05, "noise")):
        drawn.clear()
        me = types.SimpleNamespace(model=types.SimpleNamespace(conditioning_key="concat"), cond_stage_key="raw_refmap",
                                   ds=BaseDataset(size=16, transform_func="0p1tom1p1_normalizedLogarithmic_lowerbound1e-6"),
                                   noisy_observe=sigma, cond_stage_trainable=True, image_size=16, padding_mode=padding)
        ns = {"torch": _T(), "self": me, "batch": {"raw_refmap": y.clone(), "raw_refmask": mask[:, 0].clone()}, "bs": None,
              "force_c_encode": False}
        exec(textwrap.dedent("\n".join(osrc)), ns)
        blob[f"cond_{tag}"] = ns["cond"].numpy()
        blob[f"cond_{tag}_mask"] = ns["mask"].numpy()
        for i, n in enumerate(drawn):
            blob[f"cond_{tag}_noise{i}"] = n.numpy()
    np.savez_compressed(OUT / "callers_ref.npz", **blob)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    smooth_mirmap_golden()
    callers_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "small":  # only the two quick files above
        return

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
