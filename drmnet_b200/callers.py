"""Batched fast paths for the reference's callers of the renderer: one launch sequence per training
step instead of one synchronous render per (sample, BRDF vector).

* ``rendering_refmaps``  <- DRMNet.rendering_refmaps          (models/drmnet.py:667-705)
* ``synthesize_refmaps`` <- the render + normalise + transform part of DRMNet.get_input  (models/drmnet.py:523-569,
                            :610-620; dataset/basedataset.py:52-53), NaN-sentinel protocol included
* ``refmap_postprocess`` <- models/drmnet.py:610-620 alone
* ``mirmap2envmap`` / ``r0toenvmap`` <- utils/transform.py:106-144 / DRMNet.r0toenvmap (models/drmnet.py:931-941)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .renderer import B200RefMapRenderer, render_batch


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rendering_refmaps(renderer: B200RefMapRenderer, envmaps: torch.Tensor, z: torch.Tensor,
                      brdf_param_names: Optional[Sequence[str]] = None, view_from: Optional[torch.Tensor] = None,
                      new_scene: bool = False) -> torch.Tensor:
    """envmaps [B,He,We,3], z [L,B,P] -> [L,B,3,res,res]: the L BRDF vectors of sample b share envmap b and view b,
    exactly the grouping of the reference's double loop (models/drmnet.py:680-691), in one batched launch."""
    assert len(envmaps) == z.size(1)
    L, B = z.shape[0], z.shape[1]
    names = brdf_param_names or renderer.brdf_param_names
    device = envmaps.device
    if view_from is None:
        view = (renderer._new_scene_view if new_scene else renderer._view).to(device)[None].expand(B, 3)
    else:
        view = view_from.to(device)
    zz = z.to(device).transpose(0, 1).reshape(B * L, -1)  # batch-major like the reference's iteration order
    env_index = torch.arange(B, device=device).repeat_interleave(L)
    out = render_batch(envmaps, zz, view.repeat_interleave(L, dim=0), env_index=env_index, brdf_param_names=names,
                       res=renderer.refmap_res, footprint_S=renderer.footprint_S, alpha_min=renderer.alpha_min or 0.0,
                       channel_first=True)
    return out.reshape(B, L, 3, renderer.refmap_res, renderer.refmap_res).transpose(0, 1).contiguous()


def refmap_postprocess(stacks: torch.Tensor, refmap_input_scaler: Optional[float] = 0.12, transform: str = "log"
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """stacks [G,N,3,res,res] (stack 0 = LrK) -> (normalised + transformed stacks, normalizing_scale [N]).

    One fused kernel for models/drmnet.py:610-620: scale every stack of a sample so the geometric mean of LrK's
    luminance over L > 0 is ``refmap_input_scaler`` (None: no scaling), then log10(x + 0.1) + 1 (``transform="log"``,
    dataset/basedataset.py:52-53; ``None``: identity)."""
    if not stacks.is_cuda:
        raise RuntimeError("stacks must be CUDA tensors: drmnet_b200 has no CPU path")
    if transform not in ("log", None):
        raise NotImplementedError(f"transform {transform!r}: only 'log' (the shipped DRMNet configs) or None")
    stacks = stacks.contiguous().float()
    G, N, C, H, W = stacks.shape
    assert C == 3 and H == W
    out = torch.empty_like(stacks)
    scale = torch.empty(N, dtype=torch.float32, device=stacks.device)
    with torch.cuda.device(stacks.device):
        _lib.check(_lib.lib().drm_refmap_postprocess(stacks.data_ptr(), G, N, H, float(refmap_input_scaler or 0.0),
                                                     1 if transform == "log" else 0, scale.data_ptr(), out.data_ptr(),
                                                     _stream(stacks.device)))
    return out, scale


def synthesize_refmaps(renderer: B200RefMapRenderer, stacked_z: torch.Tensor, envmap: torch.Tensor,
                       view_from: torch.Tensor, cached: Optional[List[Optional[torch.Tensor]]] = None,
                       brdf_param_names: Optional[Sequence[str]] = None, refmap_input_scaler: Optional[float] = 0.12,
                       transform: str = "log"):
    """Training-data synthesis of DRMNet.get_input (models/drmnet.py:523-569, 610-620) as two launches.

    stacked_z [G,B,P] (LrK, Lrk, Lrkm1[, r0]); envmap [B,He,We,3]; view_from [B,3]; ``cached[g]`` is either None or a
    [B,3,res,res] tensor whose entries with NaN at [b,0,0,0] are cache misses (the dataset's sentinel,
    dataset/parametricrefmap.py:174-193).  Only misses are rendered.  Returns (list of G transformed [B,3,res,res]
    tensors, normalizing_scale [B], raw stacks [G,B,3,res,res])."""
    G, B = stacked_z.shape[0], stacked_z.shape[1]
    res = renderer.refmap_res
    device = envmap.device
    raw = torch.empty((G, B, 3, res, res), dtype=torch.float32, device=device)
    need = torch.ones((G, B), dtype=torch.bool, device=device)
    if cached is not None:
        for gi, c in enumerate(cached):
            if c is not None:
                raw[gi] = c.to(device)
                need[gi] = torch.isnan(raw[gi][:, 0, 0, 0])
    idx = torch.nonzero(need.T)  # (batch_idx, stack_idx): the reference's iteration order (models/drmnet.py:561)
    if idx.numel():
        b_idx, g_idx = idx[:, 0], idx[:, 1]
        out = render_batch(envmap, stacked_z.to(device)[g_idx, b_idx], view_from.to(device)[b_idx], env_index=b_idx,
                           brdf_param_names=brdf_param_names or renderer.brdf_param_names, res=res,
                           footprint_S=renderer.footprint_S, alpha_min=renderer.alpha_min or 0.0, channel_first=True)
        raw[g_idx, b_idx] = out
    post, scale = refmap_postprocess(raw, refmap_input_scaler, transform)
    return list(post), scale, raw


def mirmap2envmap(mirmap: torch.Tensor, output_shape: tuple, view=[0, 0, 1], basis: Optional[torch.Tensor] = None,
                  log_scale_interpolation: bool = False) -> torch.Tensor:
    """utils/transform.py:106-144 with its defaults: mirmap [B,C,H,W] -> envmap [B,C,OH,OW].  ``basis`` [C,H,W] divides
    the refmap first (fused form of models/drmnet.py:939)."""
    assert list(view) == [0, 0, 1], "now support [0,0,1] view direction"  # the reference's own restriction (:116)
    if log_scale_interpolation:
        raise NotImplementedError("log_scale_interpolation")
    if not mirmap.is_cuda:
        raise RuntimeError("mirmap must be a CUDA tensor: drmnet_b200 has no CPU path")
    mirmap = mirmap.contiguous().float()
    B, C, H, W = mirmap.shape
    OH, OW = output_shape
    out = torch.empty((B, C, OH, OW), dtype=torch.float32, device=mirmap.device)
    bptr = None
    if basis is not None:
        basis = basis.to(mirmap.device).contiguous().float()
        assert tuple(basis.shape) == (C, H, W)
        bptr = basis.data_ptr()
    with torch.cuda.device(mirmap.device):
        _lib.check(_lib.lib().drm_mirmap2envmap(mirmap.data_ptr(), bptr, B, C, H, W, OH, OW, out.data_ptr(),
                                                _stream(mirmap.device)))
    return out


def r0toenvmap(r0: torch.Tensor, basis_r0: torch.Tensor, envshape: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """DRMNet.r0toenvmap (models/drmnet.py:931-941): r0 [B,C,H,W] rescaled -> envmap [B,OH,OW,3]."""
    if envshape is None:
        envshape = (r0.shape[-2], r0.shape[-1] * 2)
    return mirmap2envmap(r0, envshape, basis=basis_r0).permute(0, 2, 3, 1)


def refmap_lookup(refmap: torch.Tensor, normals: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Shade pixels by their normals: colors [n,C] = bilinear lookup of refmap [B,C,H,W] (or [C,H,W]) at the (theta, phi)
    of normals [n,3]; image b owns rows offsets[b]:offsets[b+1].  The arithmetic of refmap2refimg_torch
    (utils/transform.py:170-198) for arbitrary object normals -- the stand-in for the reference's mesh path tracer in
    the parametric_img2refmap flow."""
    if not refmap.is_cuda:
        raise RuntimeError("refmap must be a CUDA tensor: drmnet_b200 has no CPU path")
    if refmap.dim() == 3:
        refmap = refmap[None]
    refmap = refmap.contiguous().float()
    normals = normals.to(refmap.device).contiguous().float()
    B, C, H, W = refmap.shape
    n = normals.shape[0]
    if offsets is None:
        assert B == 1
        offsets = torch.tensor([0, n], dtype=torch.int64, device=refmap.device)
    offsets = offsets.to(device=refmap.device, dtype=torch.int64).contiguous()
    colors = torch.empty((n, C), dtype=torch.float32, device=refmap.device)
    with torch.cuda.device(refmap.device):
        _lib.check(_lib.lib().drm_refmap_lookup(refmap.data_ptr(), normals.data_ptr(), offsets.data_ptr(), n, B, C, H, W,
                                                colors.data_ptr(), _stream(refmap.device)))
    return colors


def normalized_log_transform(x: torch.Tensor, mask: torch.Tensor, lowerbound: float = 1e-6):
    """ObsNet's conditioning transform `0p1tom1p1_normalizedLogarithmic_lowerbound1e-6` with dynamic_normalize=True
    (dataset/basedataset.py:56-76; models/obsnet.py:224,370): x [B,C,H,W], mask [B,1,H,W] or [B,H,W].
    Returns (out [B,C,H,W] in [-1,1] over the mask, (log10min [B], log10max [B]))."""
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor: drmnet_b200 has no CPU path")
    x = x.contiguous().float()
    B, C, H, W = x.shape
    mask = mask.to(x.device).float().reshape(B, H, W).contiguous()
    out = torch.empty_like(x)
    lmin = torch.empty(B, dtype=torch.float32, device=x.device)
    lmax = torch.empty(B, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().drm_normalized_log(x.data_ptr(), mask.data_ptr(), B, C, H, W, float(lowerbound),
                                                 out.data_ptr(), lmin.data_ptr(), lmax.data_ptr(), _stream(x.device)))
    return out, (lmin, lmax)
