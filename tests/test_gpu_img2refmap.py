"""GPU parity for the image -> refmap scatter: CUDA path through the C-ABI against the oracle and against the
golden outputs of the reference itself.  Bit-exact: refmask, counts, bin membership, selected pixel."""
import numpy as np
import pytest
import torch

from drmnet_b200.img2refmap import img2refmap_batch, normals_to_thetaphi, refmap_mask_make
from drmnet_b200.synth import sphere_image_inputs
from oracle.img2refmap_oracle import img2refmap_batch_oracle, img2refmap_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _run(colors, geom, res, thr, mp=0, thetaphi=False, reduce="median", offsets=None):
    if offsets is None:
        offsets = [0, len(colors)]
    off = torch.tensor(offsets, dtype=torch.int64, device=DEV)
    kw = dict(thetaphi=_gpu(geom)) if thetaphi else {}
    out = img2refmap_batch(_gpu(colors), None if thetaphi else _gpu(geom), off, res, thr, mp, reduce=reduce, **kw)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def test_sample_config0_matches_reference_output(golden_sample):
    """BASELINE config[0]: identical to the reference's refmap/refmask on data/sample (same fp32 angles)."""
    g = golden_sample
    refmap, refmask, counts, sel = _run(g["colors"], g["thetaphi_torch_cpu"], 128, float(g["thr"]), thetaphi=True)
    assert np.array_equal(refmask[0], g["refmask"])
    assert np.array_equal(refmap[0], g["refmap"])
    o = img2refmap_oracle(g["colors"], None, 128, float(g["thr"]), thetaphi=g["thetaphi_torch_cpu"])
    assert np.array_equal(counts[0], o[2]) and np.array_equal(sel[0], o[3])
    assert int(refmask.sum()) == 7621 and abs(float(refmap.astype(np.float64).sum()) - 2173.994606) < 1e-5


def test_sample_from_normals_same_device_angles(golden_sample):
    """H5: in-kernel acosf/atan2f equal torch's CUDA acos/atan2 bit for bit (the reference runs this on CUDA tensors,
    scripts/estimate.py:128-137), so bins are those of the reference run on this device."""
    g = golden_sample
    n = _gpu(g["normals"])
    ours = normals_to_thetaphi(n)
    theta = torch.acos(n[:, 1])
    phi = torch.atan2(n[:, 2], -n[:, 0] + 0.0)
    assert torch.equal(ours[:, 0], theta) and torch.equal(ours[:, 1], phi)
    refmap, refmask = refmap_mask_make(_gpu(g["colors"]), n, 128, float(g["thr"]))
    o = img2refmap_oracle(g["colors"], None, 128, float(g["thr"]), thetaphi=ours.cpu().numpy())
    assert np.array_equal(refmask.cpu().numpy(), o[1]) and np.array_equal(refmap.cpu().numpy(), o[0])
    # against the CPU-angle golden: identical on the sample (no pixel within an ulp of a cell edge)
    assert np.array_equal(refmask.cpu().numpy(), g["refmask"]) and np.array_equal(refmap.cpu().numpy(), g["refmap"])
    assert refmask.dtype == torch.bool and refmap.shape == (128, 128, 3)


@pytest.mark.parametrize("name", ["A_half_cell_res32", "B_small_window_res64", "C_overlap_res256",
                                  "D_min_points_res32", "E_nan_ties_res16", "G_nan_angles_res16",
                                  "F_wide_window_res24"])
def test_synthetic_cases(golden_synth, name):
    c = golden_synth[name]
    res, thr, mp = int(c["res"]), float(c["thr"]), int(c["min_points"])
    refmap, refmask, counts, sel = _run(c["colors"], c["thetaphi_torch_cpu"], res, thr, mp, thetaphi=True)
    o = img2refmap_oracle(c["colors"], None, res, thr, mp, thetaphi=c["thetaphi_torch_cpu"])
    for ours, ref in zip((refmap[0], refmask[0], counts[0], sel[0]), o):
        assert np.array_equal(ours, ref, equal_nan=True)
    # and against the reference's own output: mask exact, selected sum exact (ties may pick another pixel, H6)
    assert np.array_equal(refmask[0], c["refmask"])
    s_ours = (refmap[0][..., 0] + refmap[0][..., 1]) + refmap[0][..., 2]
    s_gold = (c["refmap"][..., 0] + c["refmap"][..., 1]) + c["refmap"][..., 2]
    assert np.array_equal(s_ours[refmask[0]], s_gold[refmask[0]])


def test_batched_ragged_with_empty_image_and_mean_mode(golden_synth):
    a = golden_synth["A_half_cell_res32"]
    colors, tp = a["colors"], a["thetaphi_torch_cpu"]
    offsets = [0, 0, 2500, 2500, 6000]
    for reduce in ("median", "mean"):
        ours = _run(colors, tp, 32, float(a["thr"]), 0, thetaphi=True, reduce=reduce, offsets=offsets)
        ref = img2refmap_batch_oracle(colors, None, np.array(offsets), 32, float(a["thr"]), 0, thetaphi=tp,
                                      reduce=reduce)
        for x, y in zip(ours, ref):
            assert np.array_equal(x, y)
    assert not ours[1][0].any() and not ours[1][2].any()


def test_large_sphere_image_512(golden_synth):
    """~206k masked pixels (512^2 object image), res 128: oracle equality plus size-independent properties."""
    colors, normals = sphere_image_inputs(256, seed=7)
    n = len(colors)
    assert n > 200000
    tp = normals_to_thetaphi(_gpu(normals)).cpu().numpy()
    thr = np.pi / 128 / 2
    refmap, refmask, counts, sel = _run(colors, tp, 128, thr, thetaphi=True)
    o = img2refmap_oracle(colors, None, 128, thr, thetaphi=tp)
    for ours, ref in zip((refmap[0], refmask[0], counts[0], sel[0]), o):
        assert np.array_equal(ours, ref)
    # every front-facing pixel lands in exactly one cell; every output colour is an input colour
    front = (tp[:, 1] > 0) & (tp[:, 1] < np.pi)
    assert abs(int(counts.sum()) - int(front.sum())) <= 4
    assert np.array_equal(refmap[0][refmask[0]], colors[sel[0][refmask[0]]])
    # idempotence: scattering the selected pixels again reproduces the refmap
    keep = sel[0][refmask[0]]
    again = _run(colors[keep], tp[keep], 128, thr, thetaphi=True)
    assert np.array_equal(again[0][0], refmap[0]) and np.array_equal(again[1][0], refmask[0])


def test_run_to_run_determinism():
    colors, normals = sphere_image_inputs(128, seed=3)
    outs = [_run(colors, normals, 128, np.pi / 256) for _ in range(3)]
    for o in outs[1:]:
        for x, y in zip(o, outs[0]):
            assert np.array_equal(x, y)


def test_reference_error_behaviour():
    c = torch.ones(4, 3, device=DEV)
    with pytest.raises(IndexError):
        refmap_mask_make(c[:0], c[:0], 16, 0.1)  # torch.nanmedian on an empty dim (utils/img2refmap.py:31)
    with pytest.raises(TypeError):
        refmap_mask_make(c, c, 16)  # angle_threshold=None fails at the comparison (:27)
    with pytest.raises(TypeError):
        refmap_mask_make(c.double(), c, 16, 0.1)


@pytest.mark.parametrize("res", [8, 16, 48])
def test_cell_sizes_across_both_select_paths(res):
    """Cells of ~800-1500 members (res 8: some beyond the staged select, warp-per-cell path), ~200 (res 16: staged in
    several batches per CTA) and ~20 (res 48), with duplicated colours so that keys tie inside the cells; status clean."""
    colors, normals = sphere_image_inputs(128, seed=11)
    colors = np.round(colors * 8) / 8  # many equal channel sums: the order falls back to the pixel index
    tp = normals_to_thetaphi(_gpu(normals)).cpu().numpy()
    thr = np.pi / res / 2
    out = img2refmap_batch(_gpu(colors), None, torch.tensor([0, len(colors)]), res, thr, thetaphi=_gpu(tp),
                           check_status=True)
    flags, n_big = img2refmap_batch.last_status
    assert flags == 0 and (n_big > 0) == (res == 8), (flags, n_big)
    o = img2refmap_oracle(colors, None, res, thr, thetaphi=tp)
    for ours, ref in zip(out, o):
        assert np.array_equal(ours[0].cpu().numpy(), ref)


def test_batches_beyond_the_per_call_pixel_limit_are_split(golden_synth, monkeypatch):
    import drmnet_b200.img2refmap as M
    a = golden_synth["A_half_cell_res32"]
    colors, tp = a["colors"], a["thetaphi_torch_cpu"]
    offsets = [0, 0, 2500, 2500, 4000, 6000]
    whole = _run(colors, tp, 32, float(a["thr"]), 0, thetaphi=True, offsets=offsets)
    monkeypatch.setattr(M, "MAX_PIXELS_PER_CALL", 2600)
    assert M.split_batch(offsets, 2600) == [(0, 3), (3, 4), (4, 5)]
    parts = _run(colors, tp, 32, float(a["thr"]), 0, thetaphi=True, offsets=offsets)
    for x, y in zip(whole, parts):
        assert np.array_equal(x, y)
