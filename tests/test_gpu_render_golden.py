"""Whole-image parity of the CUDA render against the fp64 oracle at the HEADLINE size (SURVEY 8c Acceptance).

tests/golden/render_cells_<He>x<We>.npz (made once by oracle/gen_render_golden.py, hours of CPU at 2000x1000) holds
the oracle's values on a strided subset of the 128x128 cells -- every 8th row / column (every 16th for the 8x8 and
16x16 footprints at 2000x1000) plus the last ones, so rows / columns 0 and 127 (the limb) are in -- for
5 synthetic envmaps x 8 BRDF vectors (z0, random zK, two schedule points, three fixed materials) x 4 views, each with
the footprint S the renderer picks for its roughness.  The inputs are regenerated here from the same seeds.

Tolerances (fp32 kernel vs fp64 oracle, identical inputs):
  relative L2 over the subset   <= 1e-4  (north_star's bound, applied to the subset instead of the whole image, which is
                                          stricter: a bright highlight usually lies between the sampled cells)
  the sharpest footprints (S = 16, the mirror end of the schedule)  <= 4e-4 on the subset; their whole-image relative
                                          L2 against the single-level GPU evaluation stays below 1e-4 (test_gpu_render)
  worst cell / image peak       <= 1e-3
"""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _check(path, max_cases=None, tol_by_S=None, loc_tol=1e-3):
    from drmnet_b200.renderer import render_batch
    from drmnet_b200.synth import synthetic_envmap
    g = np.load(path)
    He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
    meta, vals, cells = g["meta"], g["values"], g["cells"]
    n = len(meta) if max_cases is None else min(len(meta), max_cases)
    assert n >= 1
    env_seed, env = None, None
    failures, worst = [], 0.0
    for i in range(n):
        seed, zi, vi, S, nc = [int(x) for x in meta[i]]
        if seed != env_seed:
            env_seed, env = seed, torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
        z = torch.tensor(g["z"][i], dtype=torch.float32)[None]
        v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
        out = render_batch(env, z, v, res=res, footprint_S=S, alpha_min=float(g["alpha_min"]), channel_first=False,
                           check_status=True)[0].double().cpu().numpy()
        cl = cells[i][:nc]
        got, ref = out[cl[:, 0], cl[:, 1]], vals[i][:nc]
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        loc = float(np.abs(got - ref).max() / np.abs(ref).max())
        worst = max(worst, err)
        tol = (tol_by_S or {}).get(S, 4e-4 if S == 16 else 1e-4)
        if err > tol or loc > loc_tol:
            failures.append(f"env {seed} z{zi} view{vi} S={S}: rel-L2 {err:.2e} (tol {tol:.0e}) worst cell/peak {loc:.2e}")
    assert not failures, f"{len(failures)} of {n} renders out of tolerance (worst rel-L2 {worst:.2e}):\n" + "\n".join(failures)
    return n


def test_headline_size_2000x1000_vs_fp64_oracle_cells():
    n = _check(GOLDEN / "render_cells_1000x2000.npz")
    assert n >= 160  # 5 envmaps x 8 BRDF vectors x 4 views


def test_rim_bands_2000x1000_vs_fp64_oracle_cells():
    """The cells just inside the outermost ring (rows / columns 1-4 and 123-126, every 16th position along them): where
    the rim zone of the lattice passes ends and the horizon width switches -- the constants of DESIGN.md 5 were swept
    against the strided cells above, these were computed afterwards (oracle/gen_render_golden.py ... rim).

    This subset holds ONLY the hardest cells of a refmap (the lobe sits on the horizon of the normal), so its relative L2 is
    2-7 times the whole-image figure: measured 5.4e-5 (1x1), 1.3e-4 (2x2), 1.6e-4 (4x4), 2.1e-4 (8x8), 6.9e-4 (16x16),
    worst cell 2.1e-3 of the image peak -- the same with the constants before the sweep except for the 1x1 footprint
    (2.7e-5).  The bounds below pin that level."""
    n = _check(GOLDEN / "render_cells_rim_1000x2000.npz", tol_by_S={1: 1e-4, 2: 1.7e-4, 4: 2.2e-4, 8: 2.7e-4, 16: 9e-4},
               loc_tol=2.7e-3)
    assert n >= 160


def test_500x250_vs_fp64_oracle_cells():
    n = _check(GOLDEN / "render_cells_250x500.npz")
    assert n >= 160
