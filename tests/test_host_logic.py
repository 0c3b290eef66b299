"""Host-side logic of the drop-in renderer that needs no device: footprint choice, alpha clamp, BRDF parameter
composition (utils/mitsuba3_utils.py:239-243 in the reference), sensor override parsing, synthetic-input helpers that
restate the reference's dataset / schedule formulas (dataset/parametricrefmap.py:105,114-116; models/drmnet.py:481-499).
"""
import math

import numpy as np
import pytest
import torch

from drmnet_b200 import renderer as R
from drmnet_b200.synth import BRDF_PARAM_NAMES, Z0, sample_brdf, sample_view, schedule_point, envmap_directions


def _renderer(**kw):
    return R.B200RefMapRenderer(128, spp=256, denoise="simple", device="cpu", **kw)


def test_auto_footprint_is_monotone_and_follows_the_clamp():
    prev = 16
    for r in np.linspace(0.0, 1.0, 101):
        S = R.auto_footprint(float(r), 128, R.default_alpha_min(1000))
        assert S in (1, 2, 4, 8, 16) and S <= prev
        prev = S
    # the clamp of a small map removes the finest footprints: cell / alpha_min = (pi/128) / (1.25 pi/128) = 0.8 -> 4
    assert R.auto_footprint(0.0, 128, R.default_alpha_min(128)) == 4
    assert R.auto_footprint(0.0, 128, R.default_alpha_min(1000)) == 16  # (pi/128) / (1.25 pi/1000) = 6.25
    assert R.auto_footprint(0.0, 128, 1e-3) == 16
    # coarser refmaps need more sub-normals for the same lobe
    assert R.auto_footprint(0.3, 32) >= R.auto_footprint(0.3, 128) >= R.auto_footprint(0.3, 512)


def test_default_alpha_min():
    assert R.default_alpha_min(1000) == pytest.approx(1.25 * math.pi / 1000)
    assert R.default_alpha_min(10 ** 6) == 1e-3


def test_unknown_brdf_parameter_is_refused_by_name():
    with pytest.raises(NotImplementedError, match="anisotropic"):
        R._slots(["roughness.value", "anisotropic"])
    assert R._slots(BRDF_PARAM_NAMES) == [0, 1, 2, 3, 4, 5]


def test_compose_z6_clips_and_keeps_unnamed_parameters():
    r = _renderer()
    base = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    z6 = r._compose_z6(torch.tensor([1.7, -0.2]), ["roughness.value", "metallic.value"], base)
    assert z6.tolist() == pytest.approx([0.0, 0.2, 0.3, 0.4, 1.0, 0.6])
    with pytest.raises(IndexError):
        r._compose_z6(torch.tensor([0.5]), ["roughness.value", "metallic.value"], base)


def test_sensor_override_forms():
    r = _renderer()
    assert r._film_res(0) == 128

    class Film:
        def size(self):
            return (64, 32)  # (width, height) as mi.Film.size()

    class Sensor:
        def film(self):
            return Film()

    assert r._film_res(Sensor()) == 32
    assert r._film_res({"film": {"height": 48, "width": 48}}) == 48
    with pytest.raises(TypeError):
        r._film_res("front")


def test_constructor_mirrors_the_reference_attributes():
    r = _renderer(brdf_param_names=BRDF_PARAM_NAMES)
    assert r.image_size == (128, 128) and r.envmap_size == (1000, 2000) and r.spp == 256 and r.denoise == "simple"
    with pytest.raises(AssertionError):
        R.B200RefMapRenderer(128, denoise="optix", device="cpu")
    with pytest.raises(TypeError):
        _renderer().rendering(torch.zeros(6), None, envmap=None)


def test_render_batch_argument_errors_come_before_any_device_work():
    env = torch.zeros(1, 8, 16, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        R.render_batch(env, torch.zeros(1, 6), torch.tensor([[0.0, 0.0, 1.0]]))


def test_sample_brdf_and_view_follow_the_dataset():
    z = sample_brdf(3)
    g = torch.Generator().manual_seed(3)
    assert torch.equal(z, torch.rand((6,), generator=g))
    for seed in range(20):
        v = sample_view(seed)
        assert v[1] == 0 and abs(float(v.norm()) - 1) < 1e-6
        phi = math.atan2(float(v[0]), float(v[2]))
        k = (phi + math.pi) / (2 * math.pi) * 64
        assert abs(k - round(k)) < 1e-4  # one of the 64 equatorial azimuths


def test_schedule_point_matches_the_closed_form():
    zK = torch.tensor([0.2, 0.9, 0.8, 0.7, 0.4, 0.6])
    K, k, zk, zkm1 = schedule_point(zK, 0.5)
    dist = float(torch.linalg.norm(zK - torch.tensor(Z0)))
    assert K == int(math.log(0.01 / dist) / math.log(0.95)) + 2
    assert k == int(0.5 * K)
    z0 = torch.tensor(Z0)
    assert torch.allclose(zk, z0 + 0.95 ** (K - k - 1) * (zK - z0), atol=1e-6)
    assert torch.allclose(zkm1, z0 + 0.95 ** (K - k) * (zK - z0), atol=1e-6)
    # the last step lands on zK, the first one within epsilon-ish of z0
    _, _, zlast, _ = schedule_point(zK, (K - 1) / K + 1e-9)
    assert torch.allclose(zlast, zK, atol=1e-6)


def test_envmap_directions_convention():
    d = envmap_directions(4, 8)
    assert d.shape == (4, 8, 3)
    assert torch.allclose(d.norm(dim=-1), torch.ones(4, 8), atol=1e-6)
    assert d[0, :, 1].min() > 0 and d[-1, :, 1].max() < 0           # row 0 looks up (+Y)
    assert d[1, 0, 2] < 0 and d[1, 0, 0] > 0                         # column 0: just right of -Z
    assert d[1, 3, 2] > 0 or d[1, 4, 2] > 0                          # half way round: +Z


def test_gauss_legendre_nodes_nest_in_the_weight_intervals_of_coarser_orders():
    """What view_term_avg (render.cu) relies on: the m = S/Sk consecutive nodes a*m .. a*m+m-1 of the S-point rule lie
    in the a-th weight interval of the Sk-point rule, for every footprint pair the level schedule uses."""
    for S in (2, 4, 8, 16):
        xf, _ = np.polynomial.legendre.leggauss(S)
        Sk = S // 2
        while Sk >= 1:
            _, ws = np.polynomial.legendre.leggauss(Sk)
            edges = np.concatenate([[-1.0], -1.0 + np.cumsum(ws)])
            m = S // Sk
            for a in range(Sk):
                for i in range(m):
                    assert edges[a] < xf[a * m + i] < edges[a + 1], (S, Sk, a, i)
            Sk //= 2


def test_img2refmap_batches_split_at_the_per_call_pixel_limit():
    from drmnet_b200.img2refmap import MAX_PIXELS_PER_CALL, split_batch
    assert MAX_PIXELS_PER_CALL == 1 << 26  # include/drmrender.h
    assert split_batch([0, 10, 20, 30], 100) == [(0, 3)]
    assert split_batch([0, 60, 100, 150, 150, 260], 110) == [(0, 2), (2, 4), (4, 5)]
    assert split_batch([0, 0, 0], 5) == [(0, 2)]
    with pytest.raises(ValueError):
        split_batch([0, 7], 5)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): stdout is exactly one JSON line carrying
    the metric of the workload, `impl`, a `cpu_baseline` describing the run and an `e2e` object without copies."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    p = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--workload", "img2refmap",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img2refmaps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
