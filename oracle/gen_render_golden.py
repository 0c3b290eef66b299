"""Generate tests/golden/render_cells_<He>x<We>.npz: fp64 oracle values of a strided subset of refmap cells for
5 synthetic envmaps x 8 BRDF vectors x 4 views at res 128 (SURVEY 8c Acceptance).  TEST INFRASTRUCTURE ONLY.

The fp64 brute force takes minutes per sharp render at 2000x1000, so it runs once here and the values are committed;
`tests/test_gpu_render_golden.py` regenerates the same seeded inputs on the GPU box and compares the CUDA path to them.

    python -m oracle.gen_render_golden 1000 2000        # headline size (hours of CPU)
    python -m oracle.gen_render_golden 250 500
    python -m oracle.gen_render_golden 1000 2000 0 rim  # render_cells_rim_<He>x<We>.npz: the bands of cells next to the
                                                        # rim (rows / columns 1-4 and 123-126), where the rim rules switch
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from drmnet_b200.renderer import auto_footprint, default_alpha_min  # noqa: E402  (host logic only, no CUDA)
from drmnet_b200.synth import Z0, sample_brdf, sample_view, schedule_point, synthetic_envmap  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402

RES = 128
ENV_SEEDS = [1000, 1001, 1002, 1003, 1004]


def brdf_vectors(e: int):
    """8 BRDF vectors per envmap: z0, two random zK, two schedule points of the first zK, three fixed materials."""
    zK = sample_brdf(100 + e)
    _, _, z_mid, _ = schedule_point(zK, 0.6)
    _, _, z_late, _ = schedule_point(zK, 0.85)
    return [
        torch.tensor(Z0),
        zK,
        z_mid.float(),
        z_late.float(),
        sample_brdf(200 + e),
        torch.tensor([0.0, 0.8, 0.6, 0.4, 0.7, 0.5]),
        torch.tensor([0.5, 0.9, 0.5, 0.3, 0.3, 1.0]),
        torch.tensor([0.2, 0.7, 0.7, 0.9, 0.15, 0.8]),
    ]


def views():
    """three of the 64 training azimuths (dataset/parametricrefmap.py:114-116) and one general position"""
    return [sample_view(0), sample_view(1), sample_view(2), torch.tensor([0.4, 0.6, 0.7])]


def cases(He: int):
    amin = default_alpha_min(He)
    out = []
    for e, seed in enumerate(ENV_SEEDS):
        for zi, z in enumerate(brdf_vectors(e)):
            z = z.clip(0, 1)
            S = auto_footprint(float(z[4]), RES, amin)
            for vi, v in enumerate(views()):
                out.append(dict(env=e, seed=seed, zi=zi, vi=vi, z=z.numpy().astype(np.float64),
                                view=v.numpy().astype(np.float64), S=S))
    return out


def cells_for(S: int, He: int) -> np.ndarray:
    # the 8x8 and 16x16 footprints cost 64 / 256 sub-normals per cell: every 16th row / column at the headline size
    stride = 16 if (S >= 8 and He >= 1000) else 8
    return ro.strided_cells(RES, stride)


def rim_band_cells() -> np.ndarray:
    """rows / columns 1-4 and 123-126 (the cells just inside the outermost ring), every 16th position along them"""
    band = [1, 2, 3, 4, RES - 5, RES - 4, RES - 3, RES - 2]
    along = list(range(0, RES, 16)) + [RES - 1]
    cl = {(b, a) for b in band for a in along} | {(a, b) for b in band for a in along}
    return np.array(sorted(cl), np.int32)


def main():
    He, We = int(sys.argv[1]), int(sys.argv[2])
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rim = len(sys.argv) > 4 and sys.argv[4] == "rim"
    if threads:
        ro.set_threads(threads)
    out_path = ROOT / "tests" / "golden" / (f"render_cells_rim_{He}x{We}.npz" if rim else f"render_cells_{He}x{We}.npz")
    cs = cases(He)
    amin = default_alpha_min(He)
    envs = {}
    vals, cells, meta = [], [], []
    t0 = time.time()
    for n, c in enumerate(cs):
        if c["seed"] not in envs:
            envs = {c["seed"]: synthetic_envmap(He, We, seed=c["seed"])}
        cl = rim_band_cells() if rim else cells_for(c["S"], He)
        v = ro.render_oracle_cells(envs[c["seed"]], c["z"], c["view"], RES, cl, S=c["S"], alpha_min=amin)
        pad = np.full((289, 3), np.nan)
        pad[:len(v)] = v
        pc = np.full((289, 2), -1, np.int32)
        pc[:len(cl)] = cl
        vals.append(pad)
        cells.append(pc)
        meta.append([c["seed"], c["zi"], c["vi"], c["S"], len(cl)])
        print(f"[{n + 1}/{len(cs)}] seed {c['seed']} z{c['zi']} v{c['vi']} S={c['S']} cells={len(cl)} "
              f"t={time.time() - t0:.0f}s", flush=True)
        if (n + 1) % 8 == 0 or n + 1 == len(cs):
            np.savez_compressed(out_path, values=np.array(vals), cells=np.array(cells), meta=np.array(meta, np.int32),
                                z=np.array([c["z"] for c in cs[:n + 1]]), view=np.array([c["view"] for c in cs[:n + 1]]),
                                He=He, We=We, res=RES, alpha_min=amin)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
