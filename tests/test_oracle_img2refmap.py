"""Pin the img2refmap oracle against outputs of the reference itself (tests/golden, made by oracle/gen_golden.py).

The reference has no tests; the goldens are outputs of /root/reference/utils/img2refmap.py:6-37 run in the build
container on data/sample (BASELINE config[0]) and on seeded synthetic edge cases.
"""
import numpy as np
import pytest

from oracle.img2refmap_oracle import (cell_centres, img2refmap_batch_oracle, img2refmap_oracle,
                                      thetaphi_from_normals)


def _check_against_reference(case, *, use_torch_thetaphi):
    colors, normals = case["colors"], case["normals"]
    res, thr = int(case["res"]), float(case["thr"])
    mp = int(case.get("min_points", 0))
    tp = case["thetaphi_torch_cpu"] if use_torch_thetaphi else None
    refmap, refmask, counts, sel = img2refmap_oracle(colors, normals, res, thr, mp, thetaphi=tp)
    gold_map, gold_mask = case["refmap"], case["refmask"]
    # mask (= which cells are non-empty) is bit-exact
    assert np.array_equal(refmask, gold_mask)
    assert (counts[refmask] >= max(mp, 1)).all()
    # the selected SUM is bit-exact everywhere; the colour is bit-exact wherever the median sum is unique (H6)
    s_ours = (refmap[..., 0] + refmap[..., 1]) + refmap[..., 2]
    s_gold = (gold_map[..., 0] + gold_map[..., 1]) + gold_map[..., 2]
    assert np.array_equal(s_ours[refmask], s_gold[refmask])
    same = (refmap == gold_map).all(-1)
    if not same[refmask].all():
        # differing cells must be exact ties on the sum
        with np.errstate(invalid="ignore"):
            sums = (colors[:, 0] + colors[:, 1]) + colors[:, 2]
        for i, j in zip(*np.nonzero(refmask & ~same)):
            assert (sums == s_gold[i, j]).sum() > 1
    assert np.array_equal(refmap[~refmask], np.zeros_like(refmap[~refmask]))
    assert (sel[~refmask] == -1).all() and np.array_equal(colors[sel[refmask]], refmap[refmask])
    return refmap, refmask


def test_sample_config0(golden_sample):
    """BASELINE config[0]: data/sample at 128x128 -- 27 774 px, 7 621 filled cells, checksum 2173.9946."""
    refmap, refmask = _check_against_reference(golden_sample, use_torch_thetaphi=True)
    assert golden_sample["colors"].shape[0] == 27774
    assert int(refmask.sum()) == 7621
    assert abs(float(refmap.astype(np.float64).sum()) - 2173.994606) < 1e-5
    assert np.array_equal(refmap, golden_sample["refmap"])  # no ties on the sample: colours identical


def test_sample_own_angles(golden_sample):
    """Correctly-rounded angles instead of torch's SLEEF acosf/atan2f: within 2 ulp, and on the sample no pixel
    sits close enough to a cell edge for that to move it (H5): the refmap is still identical."""
    tp = thetaphi_from_normals(golden_sample["normals"])
    ref = golden_sample["thetaphi_torch_cpu"]
    ulp = np.abs(tp.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
    _check_against_reference(golden_sample, use_torch_thetaphi=False)


@pytest.mark.parametrize("name", ["A_half_cell_res32", "B_small_window_res64", "C_overlap_res256",
                                  "D_min_points_res32", "E_nan_ties_res16", "G_nan_angles_res16",
                                  "F_wide_window_res24"])
def test_synthetic_cases(golden_synth, name):
    _check_against_reference(golden_synth[name], use_torch_thetaphi=True)


def test_cell_centres_are_fp32_products():
    c = cell_centres(128)
    assert c.dtype == np.float32
    assert c[0] == np.float32(0.5) * np.float32(np.pi / 128)


def test_empty_and_bad_threshold():
    with pytest.raises(IndexError):
        img2refmap_oracle(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), 16, 0.1)
    with pytest.raises(TypeError):
        img2refmap_oracle(np.ones((4, 3), np.float32), np.ones((4, 3), np.float32), 16, None)


def test_batch_matches_single(golden_synth):
    a, e = golden_synth["A_half_cell_res32"], golden_synth["D_min_points_res32"]
    colors = np.concatenate([a["colors"][:3000], a["colors"][3000:], a["colors"][:0]])
    normals = np.concatenate([a["normals"][:3000], a["normals"][3000:], a["normals"][:0]])
    offsets = np.array([0, 3000, 6000, 6000])
    out = img2refmap_batch_oracle(colors, normals, offsets, 32, float(a["thr"]))
    one = img2refmap_oracle(a["colors"][:3000], a["normals"][:3000], 32, float(a["thr"]))
    for k in range(4):
        assert np.array_equal(out[k][0], one[k])
    assert not out[1][2].any() and (out[3][2] == -1).all()


def test_mean_mode_is_average():
    rng = np.random.default_rng(0)
    n = rng.normal(size=(500, 3)); n[:, 2] = abs(n[:, 2]); n /= np.linalg.norm(n, axis=1, keepdims=True)
    c = rng.uniform(size=(500, 3)).astype(np.float32)
    refmap, mask, counts, _ = img2refmap_oracle(c, n.astype(np.float32), 8, np.pi / 16, reduce="mean")
    tp = thetaphi_from_normals(n.astype(np.float32))
    i = np.floor(tp[:, 0] / (np.pi / 8)).astype(int); j = np.floor(tp[:, 1] / (np.pi / 8)).astype(int)
    for a, b in zip(*np.nonzero(mask)):
        m = (i == a) & (j == b)
        assert m.sum() == counts[a, b]
        np.testing.assert_allclose(refmap[a, b], c[m].mean(0), rtol=1e-5)
