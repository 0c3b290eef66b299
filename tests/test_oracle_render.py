"""Known-answer tests that pin the render oracle (parity against Mitsuba itself is UNPINNED).

K1 white furnace, K2 mirror limit against the reference's own torch renderer (golden made by oracle/gen_golden.py from
utils/transform.py:201-242), K4 linearity / azimuth equivariance / flip, K5 energy bounds.
"""
import numpy as np
import pytest

from drmnet_b200.synth import Z0, synthetic_envmap
from oracle.render_oracle import (default_alpha_min, env_records, gauss_legendre, rel_l2, render_oracle,
                                  z_from_named)

NAMES = ["metallic.value", "base_color.value.R", "base_color.value.G", "base_color.value.B", "roughness.value",
         "specular"]


def rot_azimuth(v, a):
    """Rotate about +Y so that the envmap azimuth atan2(x, -z) grows by a."""
    p = np.arctan2(v[0], -v[2]) + a
    r = np.hypot(v[0], v[2])
    return np.array([r * np.sin(p), v[1], -r * np.cos(p)])


def test_white_furnace_mirror_is_one_in_the_interior():
    """K1: env == 1, z = z0 (white metal, F = 1) -> refmap == 1 away from the limb; this is basis_r0 of
    models/drmnet.py:328-347."""
    env = np.ones((128, 256, 3), np.float32)
    r = render_oracle(env, Z0, [0, 0, 1.1], 16, S=4)
    assert np.abs(r[3:-3, 3:-3] - 1).max() < 0.01


def test_solid_angles_sum_to_4pi():
    _, E = env_records(np.ones((64, 128, 3), np.float32))
    assert abs(E[:, 0].sum() - 4 * np.pi) < 2e-3


@pytest.mark.parametrize("tag", ["v001", "v100", "vdiag"])
def test_mirror_limit_matches_reference_envmap2mirmap(golden_mirmap, tag):
    """K2: pins rows/cols/left-right, the envmap azimuth convention and the view frame.  The residual 0.15-0.2 is the
    GGX blur (alpha_min at He = 128) against the reference's box-filtered perfect mirror; any wrong convention
    (flip, transpose) gives > 0.8."""
    env, view, mir = golden_mirmap["env"], golden_mirmap[f"view_{tag}"], golden_mirmap[f"mirmap_{tag}"]
    r = render_oracle(env, Z0, view, 32, S=4)
    ok = rel_l2(r, mir)
    assert ok < 0.25
    for wrong in (r[:, ::-1], r[::-1], r.transpose(1, 0, 2)):
        assert rel_l2(wrong, mir) > 3 * ok


def test_linearity_equivariance_flip():
    env = synthetic_envmap(64, 128, seed=5)
    env2 = synthetic_envmap(64, 128, seed=6)
    z = [0.3, 0.8, 0.5, 0.2, 0.4, 0.7]
    v = np.array([0.3, 0.0, 1.0])
    a = render_oracle(env, z, v, 12, S=2)
    b = render_oracle(env2, z, v, 12, S=2)
    ab = render_oracle(2.0 * env + 0.5 * env2, z, v, 12, S=2)
    assert rel_l2(ab, 2.0 * a + 0.5 * b) < 1e-6  # fp32 rounding of the combined map
    # rolling the map by 16 of 128 columns == rotating the camera by 45 degrees about +Y
    rolled = render_oracle(np.roll(env, 16, axis=1), z, rot_azimuth(v, 2 * np.pi * 16 / 128), 12, S=2)
    assert rel_l2(rolled, a) < 1e-12
    # flip mirrors the columns (utils/mitsuba3_utils.py:38-40)
    assert rel_l2(render_oracle(env, z, v, 12, S=2, flip=True)[:, ::-1], a) < 1e-12
    # view length is irrelevant (rescaled to 1.1, :235)
    assert rel_l2(render_oracle(env, z, 3.0 * v, 12, S=2), a) < 1e-12


def test_energy_bounds_uniform_light():
    """K5: under a uniform unit environment the refmap is the directional albedo: <= 1, and for a rough dielectric
    with white base colour it stays within the range of a Disney diffuse + GGX specular surface."""
    env = np.ones((64, 128, 3), np.float32)
    z = [0.0, 1.0, 1.0, 1.0, 0.6, 0.5]
    r = render_oracle(env, z, [0, 0, 1], 12, S=2)
    interior = r[2:-2, 2:-2]
    assert interior.max() < 1.25 and interior.min() > 0.75  # Disney diffuse + specular is not energy conserving
    black = render_oracle(env, [0.0, 0.0, 0.0, 0.0, 0.6, 1.0], [0, 0, 1], 12, S=2)  # specular only, F0 = 0.08
    assert 0.03 < black[6, 6, 0] < 0.09


def test_named_parameters_and_clipping():
    z6 = z_from_named([2.0, -1.0], ["roughness.value", "metallic.value"])
    assert list(z6) == [0.0, 0.0, 0.0, 0.0, 1.0, 1.0]
    assert list(z_from_named(Z0, NAMES)) == list(Z0)
    with pytest.raises(NotImplementedError):
        z_from_named([0.5], ["clearcoat.value"])


def test_footprint_converges_to_cell_average():
    env = synthetic_envmap(64, 128, seed=9)
    z = [1.0, 1.0, 1.0, 1.0, 0.35, 1.0]
    ref = render_oracle(env, z, [0, 0, 1], 8, S=10)  # res 8: cells are 22.5 degrees wide, 3x the lobe
    errs = [rel_l2(render_oracle(env, z, [0, 0, 1], 8, S=S), ref) for S in (1, 2, 4, 6)]
    assert errs[0] > errs[1] > errs[2] > errs[3] and errs[3] < 5e-4


def test_gauss_legendre_and_alpha_min():
    x, w = gauss_legendre(4)
    assert abs(w.sum() - 1) < 1e-15 and abs((w * x ** 6).sum() - 1 / 7) < 1e-14
    assert default_alpha_min(1000) == pytest.approx(1.25 * np.pi / 1000)
    assert default_alpha_min(100000) == 1e-3
