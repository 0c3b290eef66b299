"""The C-ABI library loads without a GPU and exports every symbol include/drmrender.h declares; argument validation
and workspace queries work on the host.  No compute is launched here."""
import ctypes
import re
from pathlib import Path

import pytest

from drmnet_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not _lib.SO_PATH.exists():
        _lib.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "drmrender.h").read_text()
    declared = sorted(set(re.findall(r"\b(drm_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in drmrender.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_version_and_error_string(lib):
    assert lib.drm_version() == 200
    assert isinstance(lib.drm_last_error(), bytes)


def test_workspace_queries(lib):
    assert lib.drm_render_workspace_bytes(64, 64, 1000, 2000, 128, 1) > 0
    assert lib.drm_render_workspace_bytes(64, 64, 1000, 2000, 128, 1) > lib.drm_render_workspace_bytes(64, 1, 1000, 2000, 128, 1)
    assert lib.drm_render_workspace_bytes(1, 1, 1000, 2000, 128, 17) == 0
    assert lib.drm_render_workspace_bytes(1, 1, 1000, 2000, 128, 3) == 0  # the lattices are 1, 2, 4, 8, 16
    assert lib.drm_render_flat_workspace_bytes(1, 1, 1000, 2000, 128, 3) > 0  # the validation path takes any S <= 16
    assert lib.drm_render_workspace_bytes(0, 1, 1000, 2000, 128, 1) == 0
    a = lib.drm_img2refmap_workspace_bytes(27774, 1, 128, 3.14159 / 256)
    b = lib.drm_img2refmap_workspace_bytes(27774, 1, 128, 3.14159 / 64)
    assert 0 < a < b
    assert lib.drm_img2refmap_workspace_bytes(-1, 1, 128, 0.1) == 0


def test_argument_validation_without_a_device(lib):
    rc = lib.drm_render_refmaps(None, 1, 8, 16, None, None, None, None, 1, 8, 1, 0.0, 0, None, None, 0, None)
    assert rc == _lib.DRM_EINVAL and b"null" in lib.drm_last_error()
    rc = lib.drm_img2refmap(None, None, 0, None, 10, 1, 3, 16, 0.1, 0, 0, None, None, None, None, None, 0, None)
    assert rc == _lib.DRM_EINVAL
    rc = lib.drm_img2refmap(None, None, 0, None, 10, 1, 7, 16, 0.1, 0, 0, None, None, None, None, None, 0, None)
    assert rc == _lib.DRM_EINVAL and b"C=7" in lib.drm_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_render_options_defaults(lib):
    o = _lib.default_render_options()
    assert 0.05 <= o.kappa <= 0.2 and o.rcap > 0 and o.pixel_covariance == 1 and o.full_second_order == 1
    assert o.level_scale > 0 and o.hand_over > 0 and o.limb_boost >= 1


def test_python_wrappers_refuse_cpu_tensors():
    import torch
    from drmnet_b200.img2refmap import refmap_mask_make
    from drmnet_b200.renderer import render_batch
    with pytest.raises(RuntimeError, match="no CPU path"):
        refmap_mask_make(torch.ones(4, 3), torch.ones(4, 3), 16, 0.1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        render_batch(torch.ones(1, 8, 16, 3), torch.ones(1, 6), torch.ones(1, 3))


def test_workspace_too_small_is_reported_before_any_device_work(lib):
    """Pointers are never dereferenced on the host: a too-small workspace is refused with DRM_EWORKSPACE and a message
    that states the size needed (no CUDA call has happened yet, so this runs without a GPU)."""
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.drm_render_refmaps(p, 1, 64, 128, None, p, p, None, 1, 16, 2, 0.0, 1, p, p, 64, None)
    assert rc == _lib.DRM_EWORKSPACE
    need = lib.drm_render_workspace_bytes(1, 1, 64, 128, 16, 2)
    assert str(need).encode() in lib.drm_last_error()
    with pytest.raises(_lib.DrmError):
        _lib.check(rc)
    rc = lib.drm_img2refmap(p, p, 0, p, 100, 1, 3, 16, 0.1, 0, 0, p, p, None, None, p, 64, None)
    assert rc == _lib.DRM_EWORKSPACE
    rc = lib.drm_render_refmaps(p, 1, 64, 128, None, p, p, None, 1, 16, 17, 0.0, 1, p, p, 64, None)
    assert rc == _lib.DRM_EINVAL and b"footprint_S" in lib.drm_last_error()


def test_render_options_struct_mirrors_the_header(lib):
    """The ctypes mirror of DrmRenderOptions has the header's fields, in the header's order, with matching C types --
    a mismatch would silently shift every constant after it -- and the shipped defaults are the documented ones."""
    header = (ROOT / "include" / "drmrender.h").read_text()
    body = header[header.index("typedef struct DrmRenderOptions {"):header.index("} DrmRenderOptions;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"^\s*(const\s+int32_t\s*\*|float|int)\s*(\w+)\s*;", body, flags=re.M)
    assert len(fields) >= 20
    ctype_of = {"float": ctypes.c_float, "int": ctypes.c_int}
    mirror = _lib.RenderOptions._fields_
    assert [n for _, n in fields] == [n for n, _ in mirror]
    for (ctype, name), (_, py) in zip(fields, mirror):
        assert py is ctype_of.get(ctype, ctypes.c_void_p), name
    o = _lib.default_render_options()
    want = dict(kappa=0.1, rcap=0.06, rcap_simple=0.035, horizon=0.03, flat_scale=2.0, horizon_inner=0.06, horizon_inner_nv=4.0,
                horizon_finest=0.0375, kappa_diffuse=0.1, horizon_diffuse=0.03, level_scale=0.6, alpha_full2=0.1,
                hand_over=0.5, limb_x=4.0, limb_cells=1.3, limb_boost=2.0, limb_hand=32.0, limb_ramp=0.0)
    for k, v in want.items():
        assert getattr(o, k) == pytest.approx(v, rel=1e-6), k
    assert not o.footprint_per_render and o.collect_stats == 0
