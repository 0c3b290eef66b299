"""GPU parity for the refmap render: CUDA path through the C-ABI against the fp64 oracle on the same seeded inputs.
Tolerance: relative L2 <= 1e-4 per refmap (BASELINE.json north_star), fp32 kernel vs fp64 oracle."""
import math

import numpy as np
import pytest
import torch

from drmnet_b200.renderer import B200RefMapRenderer, render_batch
from drmnet_b200.synth import BRDF_PARAM_NAMES, Z0, sample_brdf, sample_view, schedule_point, synthetic_envmap
from oracle.render_oracle import rel_l2, render_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _render(envs, z, views, res, S, **kw):
    out = render_batch(torch.as_tensor(np.asarray(envs)).to(DEV), torch.as_tensor(np.asarray(z), dtype=torch.float32),
                       torch.as_tensor(np.asarray(views), dtype=torch.float32), res=res, footprint_S=S, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


Z_CASES = {
    "z0_mirror": list(Z0),
    "rough_dielectric": [0.0, 0.9, 0.5, 0.2, 0.7, 0.5],
    "glossy_metal": [1.0, 0.95, 0.6, 0.3, 0.3, 1.0],
    "mixed": [0.4, 0.3, 0.8, 0.6, 0.45, 0.8],
    "random7": sample_brdf(7).tolist(),
    "near_mirror_schedule": schedule_point(sample_brdf(11), 0.1)[2].tolist(),
}
VIEWS = [[0.0, 0.0, 1.1], [1.0, 0.0, 0.0], [math.sin(0.7), 0.0, math.cos(0.7)], [-0.5, 0.3, -0.8]]


@pytest.mark.parametrize("zname", list(Z_CASES))
@pytest.mark.parametrize("shape", [(64, 128), (250, 500)])
def test_parity_small_maps(zname, shape):
    """5 envmaps x 6 BRDFs x 4 views over the two map sizes (views and envmaps cycled), res 16/24, S = 2."""
    He, We = shape
    res = 16 if He == 64 else 24
    z = Z_CASES[zname]
    envs = [synthetic_envmap(He, We, seed=1000 + b) for b in range(5 if He == 64 else 2)]
    views = [VIEWS[(b + len(zname)) % 4] for b in range(len(envs))]
    ours = _render(np.stack(envs), [z] * len(envs), views, res, 2, channel_first=False)
    for b, env in enumerate(envs):
        ref = render_oracle(env, z, views[b], res, S=2)
        assert rel_l2(ours[b], ref) <= TOL, (zname, shape, b, rel_l2(ours[b], ref))


@pytest.mark.parametrize("S", [1, 3, 4, 8])
def test_parity_footprints(S):
    env = synthetic_envmap(64, 128, seed=21)
    z = Z_CASES["glossy_metal"]
    ours = _render(env[None], [z], [VIEWS[2]], 12, S, channel_first=False)[0]
    assert rel_l2(ours, render_oracle(env, z, VIEWS[2], 12, S=S)) <= TOL


def test_parity_non_tma_pitch_and_odd_sizes():
    """We*12 bytes not a multiple of 16 -> plain-load staging; sizes that are not multiples of the 32-texel tile."""
    env = synthetic_envmap(50, 101, seed=4)
    z = Z_CASES["mixed"]
    ours = _render(env[None], [z], [VIEWS[0]], 10, 2, channel_first=False)[0]
    assert rel_l2(ours, render_oracle(env, z, VIEWS[0], 10, S=2)) <= TOL


@pytest.mark.parametrize("zname", ["z0_mirror", "random7"])
def test_parity_full_size_envmap(zname):
    """BASELINE's 2000x1000 map; res 16 keeps the fp64 oracle at a few seconds (the texel sum is res independent)."""
    env = synthetic_envmap(1000, 2000, seed=1001)
    z = Z_CASES[zname]
    ours = _render(env[None], [z], [VIEWS[2]], 16, 1, channel_first=False)[0]
    assert rel_l2(ours, render_oracle(env, z, VIEWS[2], 16, S=1)) <= TOL


def test_full_size_128_properties():
    """BASELINE config[1] shape (2000x1000 -> 128x128): linearity, flip symmetry, azimuth equivariance, furnace."""
    dev_env = synthetic_envmap(1000, 2000, seed=1002, device=DEV)
    dev_env2 = synthetic_envmap(1000, 2000, seed=1003, device=DEV)
    z = torch.tensor([Z_CASES["mixed"]])
    v = torch.tensor([[0.0, 0.0, 1.0]])
    a = render_batch(dev_env[None], z, v, res=128)
    b = render_batch(dev_env2[None], z, v, res=128)
    ab = render_batch((2.0 * dev_env + 0.5 * dev_env2)[None], z, v, res=128)
    assert rel_l2(ab.cpu().numpy(), (2.0 * a + 0.5 * b).cpu().numpy()) < 2e-6
    f = render_batch(dev_env[None], z, v, res=128, flip=torch.tensor([True]))
    assert rel_l2(f.flip(-1).cpu().numpy(), a.cpu().numpy()) < 1e-6
    # rolling the map by 250 of 2000 columns == rotating the camera by 45 degrees about +Y
    ang = 2 * math.pi * 250 / 2000
    v2 = torch.tensor([[math.sin(math.pi + ang) * 1.0, 0.0, -math.cos(math.pi + ang) * 1.0]])
    r = render_batch(torch.roll(dev_env, 250, dims=1)[None], z, v2, res=128)
    assert rel_l2(r.cpu().numpy(), a.cpu().numpy()) < 2e-5
    # white furnace (basis_r0 of models/drmnet.py:328-347): 1 in the interior; toward the limb the reflected lobe is
    # compressed below the texel pitch and the texel-centre quadrature of the canonical definition degrades
    white = render_batch(torch.ones(1, 1000, 2000, 3, device=DEV), torch.tensor([list(Z0)]), v, res=128,
                         footprint_S=2)[0]
    assert (white[:, 32:-32, 32:-32] - 1).abs().max() < 2e-3
    assert (white[:, 16:-16, 16:-16] - 1).abs().max() < 5e-2
    assert a.shape == (1, 3, 128, 128) and torch.isfinite(a).all()


def test_batch_env_index_layouts_and_splits():
    """G renders sharing one envmap (models/drmnet.py:561-569 groups), channel-first vs channel-last, and the
    split-texel path (N = 1) against the unsplit path (N large)."""
    envs = np.stack([synthetic_envmap(64, 128, seed=s) for s in (31, 32)])
    zs = [Z_CASES["z0_mirror"], Z_CASES["mixed"], Z_CASES["rough_dielectric"]] * 14  # N = 42 -> no splits
    idx = torch.tensor([i % 2 for i in range(len(zs))], dtype=torch.int32)
    views = [VIEWS[i % 4] for i in range(len(zs))]
    cf = _render(envs, zs, views, 16, 1, env_index=idx, channel_first=True)
    cl = _render(envs, zs, views, 16, 1, env_index=idx, channel_first=False)
    assert np.array_equal(cf.transpose(0, 2, 3, 1), cl)
    for i in (0, 1, 2, 5):
        one = _render(envs[int(idx[i]):int(idx[i]) + 1], [zs[i]], [views[i]], 16, 1, channel_first=False)[0]  # splits
        assert rel_l2(one, cl[i]) < 5e-6
        assert rel_l2(cl[i], render_oracle(envs[int(idx[i])], zs[i], views[i], 16, S=1)) <= TOL


def test_determinism():
    env = synthetic_envmap(64, 128, seed=77)
    outs = [_render(env[None], [Z_CASES["mixed"]], [VIEWS[3]], 16, 2) for _ in range(3)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_dropin_class_statefulness_and_attributes():
    """Same ctor kwargs/attributes as MitsubaRefMapRenderer (utils/mitsuba3_utils.py:318-339) and the 'None keeps the
    last value' protocol used by models/drmnet.py:561-569."""
    r = B200RefMapRenderer(refmap_res=16, spp=256, envmap_size=(64, 128), denoise="simple",
                           brdf_param_names=BRDF_PARAM_NAMES, footprint_S=2)
    assert (r.refmap_res, r.image_size, r.spp, r.envmap_size, r.denoise) == (16, (16, 16), 256, (64, 128), "simple")
    env = torch.from_numpy(synthetic_envmap(64, 128, seed=8)).to(DEV)
    z = torch.tensor(Z_CASES["mixed"])
    view = torch.tensor(VIEWS[2])
    a = r.rendering(z, BRDF_PARAM_NAMES, envmap=env, view_from=view, channel_first=True)
    assert a.shape == (3, 16, 16) and a.is_cuda and a.dtype == torch.float32
    b = r.rendering(z, None, channel_first=False)  # keeps envmap and view; names fall back to the ctor list
    assert torch.equal(a.permute(1, 2, 0), b)
    ref = render_oracle(env.cpu().numpy(), Z_CASES["mixed"], VIEWS[2], 16, S=2)
    assert rel_l2(b.cpu().numpy(), ref) <= TOL
    # unnamed parameters keep their last value in the persistent scene
    c = r.rendering(torch.tensor([0.9]), ["roughness.value"])
    z2 = list(Z_CASES["mixed"]); z2[4] = 0.9
    assert rel_l2(c.cpu().numpy(), render_oracle(env.cpu().numpy(), z2, VIEWS[2], 16, S=2)) <= TOL
    # new_scene renders with fresh BSDF defaults and does not disturb the persistent scene
    d = r.rendering(torch.tensor([0.9]), ["roughness.value"], envmap=env, new_scene=True)
    assert rel_l2(d.cpu().numpy(), render_oracle(env.cpu().numpy(), [0.9], [0, 0, 1.1], 16, S=2,
                                                 names=["roughness.value"])) <= TOL
    with pytest.raises(AssertionError):
        r.rendering(z, BRDF_PARAM_NAMES, envmap=env[0])  # envmap must be 3-D (utils/mitsuba3_utils.py:424)
    bad = env.clone(); bad[0, 0, 0] = float("nan")
    with pytest.raises(AssertionError):
        r.rendering(z, BRDF_PARAM_NAMES, envmap=bad)


# ---- footprint hierarchy (coarser lattices for texel tiles far from the lobe) --------------------------------------
def _with_env(name, flag, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = flag
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


def _with_levels(flag, fn):
    return _with_env("DRM_RENDER_LEVELS", flag, fn)


@pytest.mark.parametrize("case", [("z0_mirror", 16), ("near_mirror_schedule", 16), ("glossy_metal", 4), ("mixed", 2)])
def test_hierarchy_matches_single_level_full_size(case):
    """BASELINE size (2000x1000 -> 128x128): the level schedule against the single-level evaluation of the same
    canonical sum (every sub-normal x every texel), all 16 384 pixels."""
    zname, S = case
    env = synthetic_envmap(1000, 2000, seed=1004, device=DEV)[None]
    z = torch.tensor([Z_CASES[zname]])
    v = torch.tensor([VIEWS[2]])
    hier = _with_levels("1", lambda: render_batch(env, z, v, res=128, footprint_S=S))
    flat = _with_levels("0", lambda: render_batch(env, z, v, res=128, footprint_S=S))
    torch.cuda.synchronize()
    assert rel_l2(hier.cpu().numpy(), flat.cpu().numpy()) <= 7e-5


@pytest.mark.parametrize("zname", ["z0_mirror", "glossy_metal"])
def test_hierarchy_parity_vs_oracle_windows(zname):
    """S = 16 at res 128 on a 1000x500 map against the fp64 oracle, on two 4x4-cell windows: the brightest cell of the
    refmap (a light source in the lobe) and a dim one (tail contributions only)."""
    env = synthetic_envmap(500, 1000, seed=1004)
    z = Z_CASES[zname]
    ours = _render(env[None], [z], [VIEWS[0]], 128, 16, channel_first=False)[0]
    lum = ours.sum(-1)
    bi, bj = np.unravel_index(np.argmax(lum[8:-8, 8:-8]), lum[8:-8, 8:-8].shape)
    bi, bj = int(bi) + 8, int(bj) + 8
    for (i0, j0) in ((bi - 2, bj - 2), (60, 20)):
        win = (i0, i0 + 4, j0, j0 + 4)
        ref = render_oracle(env, z, VIEWS[0], 128, S=16, window=win)
        a, b = ours[win[0]:win[1], win[2]:win[3]], ref[win[0]:win[1], win[2]:win[3]]
        assert rel_l2(a, b) <= TOL, (zname, win, rel_l2(a, b))


def test_limb_window_vs_oracle_local_accuracy():
    """The limb columns of the refmap (n.v -> 0) are where the accelerated stages are least accurate *locally*: the
    single-level evaluation matches the fp64 oracle to 1e-6 there, the level / coarse-map schedule to ~1e-3 on a forced
    S = 16 glossy render (scripts/limb_probe.py).  The window carries ~1e-3 of the image norm, so the
    whole-image error stays below 1e-4; this test pins both numbers so the local error cannot grow unnoticed."""
    env = synthetic_envmap(500, 1000, seed=1004)
    z = Z_CASES["glossy_metal"]
    win = (62, 66, 124, 128)
    ref = render_oracle(env, z, VIEWS[0], 128, S=16, window=win)[win[0]:win[1], win[2]:win[3]]
    hier = _render(env[None], [z], [VIEWS[0]], 128, 16, channel_first=False)[0]
    flat = _with_env("DRM_RENDER_COARSE", "0", lambda: _with_levels(
        "0", lambda: _render(env[None], [z], [VIEWS[0]], 128, 16, channel_first=False)))[0]
    assert rel_l2(flat[win[0]:win[1], win[2]:win[3]], ref) <= 1e-5
    assert rel_l2(hier[win[0]:win[1], win[2]:win[3]], ref) <= 2.5e-3
    assert rel_l2(hier, flat) <= 7e-5


# ---- coarse-map routes (diffuse lobe / very rough specular lobe gathered from the 4x4 energy-centroid map) ----------
Z_CASES["moderately_rough"] = [0.5, 0.9, 0.8, 0.3, 0.62, 0.7]


@pytest.mark.parametrize("zname,S", [("rough_dielectric", 1), ("mixed", 2), ("random7", 2), ("moderately_rough", 1)])
def test_coarse_routes_match_raw_map_full_size(zname, S):
    """2000x1000 -> 128x128, all pixels: routes through the coarse map against everything gathered from the raw map."""
    env = synthetic_envmap(1000, 2000, seed=1004, device=DEV)[None]
    z = torch.tensor([Z_CASES[zname]])
    v = torch.tensor([VIEWS[1]])
    fast = render_batch(env, z, v, res=128, footprint_S=S)
    full = _with_env("DRM_RENDER_COARSE", "0", lambda: _with_levels("0", lambda: render_batch(env, z, v, res=128, footprint_S=S)))
    torch.cuda.synchronize()
    assert rel_l2(fast.cpu().numpy(), full.cpu().numpy()) <= 7e-5


@pytest.mark.parametrize("z", [[0.2, 0.9, 0.8, 0.7, 0.85, 0.6], [0.0, 1.0, 1.0, 1.0, 0.3, 0.0], [1.0, 0.9, 0.5, 0.3, 1.0, 1.0],
                               [0.5, 0.9, 0.8, 0.3, 0.62, 0.7], [1.0, 1.0, 0.9, 0.8, 0.6, 1.0]])
def test_coarse_routes_parity_vs_oracle_full_size(z):
    """Very rough (both lobes from the 4x4 map), diffuse-dominated (diffuse from the 4x4 map), rough metal, and two
    moderately rough cases (both lobes from the 2x2 map)."""
    env = synthetic_envmap(1000, 2000, seed=1007)
    ours = _render(env[None], [z], [VIEWS[3]], 16, 1, channel_first=False)[0]
    assert rel_l2(ours, render_oracle(env, z, VIEWS[3], 16, S=1)) <= TOL


@pytest.mark.parametrize("res,S", [(24, 8), (40, 16), (17, 4)])
def test_hierarchy_partial_blocks_and_far_near_pair(res, S):
    """Refmap sizes that are not multiples of the far launch's 16-cell blocks (or of the CTA's cell block): the far/near
    pair and the level schedule against the single-level evaluation, and against the oracle on a window."""
    env = synthetic_envmap(250, 500, seed=1004)
    z = Z_CASES["z0_mirror"] if S > 4 else Z_CASES["glossy_metal"]
    hier = _render(env[None], [z], [VIEWS[2]], res, S, channel_first=False)[0]
    flat = _with_levels("0", lambda: _render(env[None], [z], [VIEWS[2]], res, S, channel_first=False))[0]
    assert rel_l2(hier, flat) <= 7e-5
    win = (res // 2 - 1, res // 2 + 1, res - 3, res)  # includes the last, partial block column
    ref = render_oracle(env, z, VIEWS[2], res, S=S, window=win)
    a, b = hier[win[0]:win[1], win[2]:win[3]], ref[win[0]:win[1], win[2]:win[3]]
    assert rel_l2(a, b) <= TOL


# ---- regressions found by the seed sweeps of scripts/bisect_probe.py and scripts/scale_probe.py ---------------------
def test_limb_cells_random_renders_full_size():
    """Random sharp and glossy renders (the schedule's own footprints) on 2000x1000 maps: every accelerated stage against
    the un-accelerated sum.  Before the coarse lattices carried the mean view term G1(n.v)/(4 n.v) of their sub-region
    these were off by up to 3e-4, all of it in the limb rows / columns of the refmap."""
    from drmnet_b200.synth import sample_brdf, sample_view
    picks = [1, 5, 13, 21]
    env = torch.stack([synthetic_envmap(1000, 2000, seed=100 + (i % 8), device=DEV) for i in picks])
    z = torch.stack([sample_brdf(500 + i) for i in picks])
    v = torch.stack([sample_view(500 + i) for i in picks])
    idx = torch.arange(len(picks), dtype=torch.int32)
    fast = render_batch(env, z, v, env_index=idx, res=128, footprint_S=None)
    full = _with_env("DRM_RENDER_COARSE", "0", lambda: _with_levels(
        "0", lambda: render_batch(env, z, v, env_index=idx, res=128, footprint_S=None)))
    torch.cuda.synchronize()
    for k in range(len(picks)):
        e = rel_l2(fast[k].cpu().numpy(), full[k].cpu().numpy())
        assert e <= 8e-5, (picks[k], e)
    # the limb columns on their own (they carry < 1 % of the image norm)
    a, b = fast[..., [0, 127]].cpu().numpy(), full[..., [0, 127]].cpu().numpy()
    assert rel_l2(a, b) <= 2e-3


@pytest.mark.parametrize("scale", ["2.0", "10.0"])
def test_near_mode_far_coarse_partition_with_wide_near_field(scale):
    """A near field (tk2[0]) wider than the distance at which the 2x2 map takes over: the raw-map and coarse-map far
    launches must both yield to the per-cell kernel there (pairs were counted twice at the pole cell before)."""
    env = synthetic_envmap(250, 500, seed=1004, device=DEV)[None]
    z = torch.tensor([Z_CASES["z0_mirror"]])
    v = torch.tensor([VIEWS[2]])
    hier = _with_env("DRM_RENDER_LEVEL_SCALE", scale, lambda: render_batch(env, z, v, res=128, footprint_S=16))
    flat = _with_levels("0", lambda: render_batch(env, z, v, res=128, footprint_S=16))
    torch.cuda.synchronize()
    assert rel_l2(hier.cpu().numpy(), flat.cpu().numpy()) <= 1e-5


def test_dropin_aovs_and_sensor_override():
    """return_normal / return_depth outputs (utils/mitsuba3_utils.py:196-214) and the `sensor=` film-size override used by
    DRMNet.instantiate_brdf_model (models/drmnet.py:331-343)."""
    class Film:
        def size(self):
            return (24, 24)

    class Sensor:  # stands in for the mitsuba sensor object: only its film size is consulted
        def film(self):
            return Film()

    r = B200RefMapRenderer(refmap_res=16, envmap_size=(64, 128), return_normal=True, return_depth=True,
                           brdf_param_names=BRDF_PARAM_NAMES, footprint_S=1)
    env = torch.ones(64, 128, 3)  # a CPU tensor, as instantiate_brdf_model passes (the reference calls .cuda() on it)
    img, normal, depth = r.rendering(torch.tensor(Z0), BRDF_PARAM_NAMES, envmap=env, channel_first=True)
    assert img.shape == (3, 16, 16) and normal.shape == (3, 16, 16) and depth.shape == (1, 16, 16)
    assert torch.allclose(normal.norm(dim=0), torch.ones(16, 16, device=DEV), atol=1e-6)
    assert normal[2].min() > 0 and abs(float(normal[1, 0, 8]) - math.cos(0.5 * math.pi / 16)) < 1e-6  # row 0 faces +up
    assert float(depth.min()) > 0.09 and float(depth.max()) < 1.1
    big = r.rendering(torch.tensor(Z0), BRDF_PARAM_NAMES, sensor=Sensor())[0]
    assert big.shape == (24, 24, 3)
    assert (big[6:-6, 6:-6] - 1).abs().max() < 0.05  # white furnace, z0
    big2 = r.rendering(torch.tensor(Z0), BRDF_PARAM_NAMES, sensor={"film": {"height": 20, "width": 20}})[0]
    assert big2.shape == (20, 20, 3)
