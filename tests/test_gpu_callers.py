"""GPU parity for the 'next' rows: fused post-processing (N1), batched caller fast paths (N1), mirmap2envmap (N2)."""
import numpy as np
import pytest
import torch

from drmnet_b200.callers import (mirmap2envmap, r0toenvmap, refmap_postprocess, rendering_refmaps, synthesize_refmaps)
from drmnet_b200.renderer import B200RefMapRenderer
from drmnet_b200.synth import BRDF_PARAM_NAMES, Z0, sample_brdf, sample_view, schedule_point, synthetic_envmap
from oracle.callers_oracle import mirmap2envmap_oracle, postprocess_oracle
from oracle.render_oracle import rel_l2

GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_postprocess_matches_oracle():
    g = torch.Generator().manual_seed(1)
    stacks = torch.exp(torch.randn(4, 5, 3, 128, 128, generator=g))
    stacks[0, 2, :, :10] = 0
    out, scale = refmap_postprocess(stacks.to(DEV))
    ref, s = postprocess_oracle(stacks.numpy())
    assert np.allclose(scale.cpu().numpy(), s, rtol=2e-6)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-6
    raw, one = refmap_postprocess(stacks.to(DEV), None, None)
    assert torch.equal(raw.cpu(), stacks) and torch.all(one == 1)


def test_mirmap2envmap_matches_reference_output_and_oracle(golden_mirmap):
    mir = torch.from_numpy(golden_mirmap["mirmap_v001"]).permute(2, 0, 1)[None].to(DEV)
    env = mirmap2envmap(mir, (32, 64))[0].permute(1, 2, 0).cpu().numpy()
    ref = golden_mirmap["envmap_from_mirmap_v001"]  # the reference's own output
    assert np.abs(env - ref).max() <= 2e-5 * np.abs(ref).max()
    basis = torch.rand(3, 32, 32, device=DEV) + 0.5
    big = r0toenvmap(mir.repeat(3, 1, 1, 1), basis, (128, 256))
    assert big.shape == (3, 128, 256, 3)
    o = mirmap2envmap_oracle(mir.cpu().numpy(), (128, 256), basis=basis.cpu().numpy())[0].transpose(1, 2, 0)
    assert np.abs(big[1].cpu().numpy() - o).max() <= 2e-5 * np.abs(o).max()


def test_batched_callers_equal_the_reference_loop():
    """rendering_refmaps / synthesize_refmaps against the reference's per-render loop structure
    (models/drmnet.py:561-569, 680-691) driven through the drop-in class."""
    B, res = 3, 16
    envs = torch.stack([torch.from_numpy(synthetic_envmap(64, 128, seed=40 + b)) for b in range(B)]).to(DEV)
    views = torch.stack([sample_view(40 + b) for b in range(B)])
    zK = torch.stack([sample_brdf(40 + b) for b in range(B)])
    sched = [schedule_point(zK[b], 0.3) for b in range(B)]
    zk = torch.stack([s[2] for s in sched]).float()
    zkm1 = torch.stack([s[3] for s in sched]).float()
    z0 = torch.tensor(Z0).expand(B, 6)
    stacked_z = torch.stack([zK, zk, zkm1, z0])  # [G,B,P]
    r = B200RefMapRenderer(refmap_res=res, spp=256, envmap_size=(64, 128), denoise="simple",
                           brdf_param_names=BRDF_PARAM_NAMES, footprint_S=2)
    # the reference loop: envmap/view passed on the first render of each sample only
    loop = torch.empty(4, B, 3, res, res, device=DEV)
    for b in range(B):
        for gi in range(4):
            loop[gi, b] = r.rendering(stacked_z[gi, b], BRDF_PARAM_NAMES, envmap=envs[b] if gi == 0 else None,
                                      view_from=views[b] if gi == 0 else None, channel_first=True)
    batched = rendering_refmaps(r, envs, stacked_z, view_from=views)
    assert rel_l2(batched.cpu().numpy(), loop.cpu().numpy()) < 5e-6

    # NaN-sentinel protocol: sample 1 has LrK cached, everything else is a miss
    cached_K = torch.full((B, 3, res, res), float("nan"), device=DEV)
    cached_K[1] = loop[0, 1]
    outs, scale, raw = synthesize_refmaps(r, stacked_z, envs, views, cached=[cached_K, None, None, None])
    assert torch.equal(raw[0, 1], loop[0, 1]) and rel_l2(raw.cpu().numpy(), loop.cpu().numpy()) < 5e-6
    ref, s = postprocess_oracle(loop.cpu().numpy())
    assert np.allclose(scale.cpu().numpy(), s, rtol=1e-5)
    for gi in range(4):
        assert np.abs(outs[gi].cpu().numpy() - ref[gi]).max() < 1e-4


def test_refmap_lookup_matches_reference_and_oracle(golden_mirmap):
    from drmnet_b200.callers import refmap_lookup
    from oracle.callers_oracle import refmap_lookup_oracle
    from drmnet_b200.synth import sphere_normals
    n, mask = sphere_normals(24)
    refmap = torch.from_numpy(golden_mirmap["refimg_refmap"]).to(DEV)
    ours = refmap_lookup(refmap, torch.from_numpy(n[mask]).to(DEV)).cpu().numpy()
    ref = golden_mirmap["refimg_image"][:, mask].T  # the reference's refmap2refimg_torch
    assert np.abs(ours - ref).max() <= 3e-5 * np.abs(ref).max()
    # batched, arbitrary normals
    g = torch.Generator().manual_seed(5)
    maps = torch.exp(torch.randn(3, 3, 64, 64, generator=g)).to(DEV)
    nrm = torch.nn.functional.normalize(torch.randn(5000, 3, generator=g), dim=-1)
    offsets = torch.tensor([0, 1000, 1000, 5000])
    out = refmap_lookup(maps, nrm, offsets).cpu().numpy()
    for b, (lo, hi) in enumerate([(0, 1000), (1000, 1000), (1000, 5000)]):
        if hi > lo:
            o = refmap_lookup_oracle(maps[b].cpu().numpy(), nrm[lo:hi].numpy())
            assert np.abs(out[lo:hi] - o).max() <= 3e-5 * np.abs(o).max()


def test_normalized_log_transform_matches_oracle():
    from drmnet_b200.callers import normalized_log_transform
    from oracle.callers_oracle import normalized_log_oracle
    g = torch.Generator().manual_seed(6)
    x = torch.exp(torch.randn(4, 3, 128, 128, generator=g) * 2)
    x[1, :, :5] = 0.0
    mask = (torch.rand(4, 1, 128, 128, generator=g) > 0.4)
    out, (lmin, lmax) = normalized_log_transform(x.to(DEV), mask.to(DEV))
    ref, rmin, rmax = normalized_log_oracle(x.numpy(), mask.numpy())
    assert np.allclose(lmin.cpu().numpy(), rmin, atol=2e-6) and np.allclose(lmax.cpu().numpy(), rmax, atol=2e-6)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5


def test_config3_pipeline_render_shade_scatter_round_trip():
    """BASELINE config[2] data flow (parametricrefmap + parametric_img2refmap): render a refmap, shade a 256^2 sphere image
    with it (N3 stand-in for the mesh renderer), scatter the image back into a refmap (K2).  Every filled cell must hold a
    bilinear sample of the rendered refmap taken inside that cell: within the cell's neighbourhood range."""
    from drmnet_b200.callers import refmap_lookup
    from drmnet_b200.img2refmap import img2refmap_batch
    from drmnet_b200.renderer import render_batch
    from drmnet_b200.synth import sphere_normals
    res = 128
    env = torch.from_numpy(synthetic_envmap(250, 500, seed=9)).to(DEV)
    z = torch.tensor([[0.3, 0.8, 0.6, 0.4, 0.5, 0.7]])
    refmap = render_batch(env[None], z, torch.tensor([[0.0, 0.0, 1.0]]), res=res, footprint_S=1)  # [1,3,res,res]
    n, mask = sphere_normals(128)
    normals = torch.from_numpy(n[mask]).to(DEV)
    colors = refmap_lookup(refmap, normals)
    offsets = torch.tensor([0, normals.shape[0]], dtype=torch.int64, device=DEV)
    back, filled, counts, _ = img2refmap_batch(colors, normals, offsets, res, float(np.pi / res / 2))
    back, filled, src = back[0].cpu().numpy(), filled[0].cpu().numpy(), refmap[0].permute(1, 2, 0).cpu().numpy()
    assert filled.sum() > 0.6 * res * res
    pad = np.pad(src, ((1, 1), (1, 1), (0, 0)), mode="edge")
    lo = np.minimum.reduce([pad[1 + di:res + 1 + di, 1 + dj:res + 1 + dj] for di in (-1, 0, 1) for dj in (-1, 0, 1)])
    hi = np.maximum.reduce([pad[1 + di:res + 1 + di, 1 + dj:res + 1 + dj] for di in (-1, 0, 1) for dj in (-1, 0, 1)])
    ok = (back >= lo * (1 - 1e-5)) & (back <= hi * (1 + 1e-5))
    assert ok[filled].all()
    assert rel_l2(back[filled], src[filled]) < 0.05


def test_round_trip_render_then_mirmap2envmap_recovers_the_envmap():
    """r0toenvmap(render(z0)) ~ envmap on the hemisphere the mirror sphere shows sharply.  Ties the renderer's
    geometry (rows, columns, left/right, envmap azimuth) to the reference's mirmap2envmap, whose kernel is pinned by a
    golden of the reference's own output."""
    from drmnet_b200.callers import r0toenvmap
    from drmnet_b200.renderer import render_batch
    from drmnet_b200.synth import Z0, envmap_directions
    g = torch.Generator().manual_seed(11)
    low = torch.exp(0.8 * torch.randn(1, 3, 8, 16, generator=g))
    low = torch.cat([low, low[..., :1]], -1)  # periodic in azimuth
    env = torch.nn.functional.interpolate(low, size=(128, 257), mode="bicubic", align_corners=True)[0, :, :, :256]
    env = env.clamp_min(0.05).permute(1, 2, 0).contiguous().to(DEV)  # smooth HDR-ish map [128,256,3]
    z0 = torch.tensor([list(Z0)])
    view = torch.tensor([[0.0, 0.0, 1.0]])
    r0 = render_batch(env[None], z0, view, res=128, footprint_S=4)
    basis = render_batch(torch.ones_like(env)[None], z0, view, res=128, footprint_S=4)[0]
    rec = r0toenvmap(r0, basis, (128, 256))[0]  # [128,256,3]
    d = envmap_directions(128, 256, device=DEV)
    front = d[..., 2] > 0.3  # reflected off normals within ~50 degrees of the viewer: well away from the limb
    ok = rel_l2(rec[front].cpu().numpy(), env[front].cpu().numpy())
    assert ok < 0.1, ok
    for wrong in (env.flip(1), env.flip(0), torch.roll(env, 64, dims=1)):
        assert rel_l2(rec[front].cpu().numpy(), wrong[front].cpu().numpy()) > 3 * ok


def test_obsnet_condition_matches_the_reference_lines():
    """N4: the fused conditioning kernel and the fixed-parameter transform / rescale against goldens produced by the
    reference's own code (oracle/gen_golden.py callers_golden: models/obsnet.py:663-695, dataset/basedataset.py:56-110),
    and against the fp64 oracle on a refmap-sized batch."""
    from drmnet_b200.callers import normalized_log_apply, normalized_log_rescale, obsnet_condition
    from oracle.callers_oracle import obsnet_condition_oracle
    g = np.load(GOLDEN / "callers_ref.npz")
    raw, mask = torch.from_numpy(g["nlog_in"]).to(DEV), torch.from_numpy(g["nlog_mask"][:, 0]).to(DEV)
    cond, m, (lmin, lmax) = obsnet_condition(raw, mask)
    assert np.allclose(cond.cpu().numpy(), g["cond_plain"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(m.cpu().numpy(), g["cond_plain_mask"]) and m.shape == (4, 1, 16, 16)
    assert np.allclose(lmin.cpu().numpy(), g["nlog_min"], atol=2e-6) and np.allclose(lmax.cpu().numpy(), g["nlog_max"], atol=2e-6)
    cond, _, _ = obsnet_condition(raw, mask, noisy_observe=0.05, observe_noise=torch.from_numpy(g["cond_noisy_noise0"]),
                                  padding_mode="noise", padding_noise=torch.from_numpy(g["cond_noisy_noise1"]))
    assert np.allclose(cond.cpu().numpy(), g["cond_noisy"], rtol=2e-5, atol=2e-5)
    # LrK with the raw refmap's parameters (models/obsnet.py:371), and the way back
    t = normalized_log_apply(torch.from_numpy(g["nlog_fixed_in"]).to(DEV), (lmin, lmax))
    assert np.allclose(t.cpu().numpy(), g["nlog_fixed_out"], rtol=2e-5, atol=2e-5)
    w = torch.from_numpy(g["nlog_rescale_in"]).to(DEV)
    assert np.allclose(normalized_log_rescale(w, (lmin, lmax)).cpu().numpy(), g["nlog_rescale_out"], rtol=3e-5)
    assert np.allclose(normalized_log_rescale(w, (lmin, lmax), 0.5).cpu().numpy(), g["nlog_rescale_clamped_out"], rtol=3e-5)
    back = normalized_log_rescale(t, (lmin, lmax))
    assert np.allclose(back.cpu().numpy(), np.clip(g["nlog_fixed_in"], 1e-6, None), rtol=1e-4)
    # refmap-sized batch against the oracle; drawn-here noise follows torch's generator
    gen = torch.Generator().manual_seed(8)
    x = torch.exp(torch.randn(6, 3, 128, 128, generator=gen) * 2)
    mk = torch.rand(6, 128, 128, generator=gen) > 0.5
    n1, n2 = torch.randn(6, 3, 128, 128, generator=gen), torch.randn(6, 3, 128, 128, generator=gen)
    cond, _, _ = obsnet_condition(x.to(DEV), mk.to(DEV), noisy_observe=0.1, observe_noise=n1, padding_mode="noise",
                                  padding_noise=n2)
    ref, _, _ = obsnet_condition_oracle(x.numpy(), mk.numpy(), noisy_observe=0.1, observe_noise=n1.numpy(),
                                        padding_noise=n2.numpy())
    assert np.abs(cond.cpu().numpy() - ref).max() < 3e-5
    torch.manual_seed(3)
    a = obsnet_condition(x.to(DEV), mk.to(DEV), noisy_observe=0.1, padding_mode="noise")[0]
    torch.manual_seed(3)
    e1 = torch.randn_like(x.to(DEV)); e2 = torch.randn_like(x.to(DEV))
    b = obsnet_condition(x.to(DEV), mk.to(DEV), noisy_observe=0.1, observe_noise=e1, padding_mode="noise", padding_noise=e2)[0]
    assert torch.equal(a, b)
    with pytest.raises(NotImplementedError):
        obsnet_condition(x.to(DEV), mk.to(DEV), padding_mode="reflect")
    with pytest.raises(AssertionError):
        normalized_log_apply(x.to(DEV), (lmin, lmax))


def test_rendering_refmaps_list_inputs_partial_names_and_state():
    """`[envmap]` / `[view_from]` lists as models/drmnet.py:943-952 passes them, a partial parameter list whose
    unnamed slots come from the persistent scene, and the renderer's state after the call -- all equal to the loop of
    stateful `rendering` calls that DRMNet.rendering_refmaps runs (models/drmnet.py:694-703)."""
    res = 16
    env = torch.from_numpy(synthetic_envmap(64, 128, seed=71)).to(DEV)
    view = sample_view(71)

    def fresh():
        r = B200RefMapRenderer(refmap_res=res, spp=256, envmap_size=(64, 128), denoise="simple",
                               brdf_param_names=BRDF_PARAM_NAMES, footprint_S=2)
        # a first render leaves non-default metallic / specular in the persistent scene
        r.rendering(torch.tensor([0.7, 0.6, 0.5, 0.4, 0.35, 0.6]), BRDF_PARAM_NAMES, envmap=env, channel_first=True)
        return r

    names = ["base_color.value.R", "base_color.value.G", "base_color.value.B", "roughness.value"]
    z = torch.tensor([[[0.9, 0.8, 0.7, 0.3]], [[0.2, 0.3, 0.4, 0.6]]])  # [L=2, B=1, 4]
    a, b = fresh(), fresh()
    loop = torch.stack([a.rendering(z[i, 0], names, envmap=env if i == 0 else None, view_from=view if i == 0 else None,
                                    channel_first=True) for i in range(2)])[:, None]
    batched = rendering_refmaps(b, [env], z, brdf_param_names=names, view_from=[view])
    assert batched.shape == (2, 1, 3, res, res)
    assert rel_l2(batched.cpu().numpy(), loop.cpu().numpy()) < 5e-6
    assert torch.equal(a._bsdf.cpu(), b._bsdf.cpu()) and torch.equal(a._view, b._view) and torch.equal(a._envmap, b._envmap)
    # and the next stateful render continues from the same scene
    nxt = torch.tensor([0.5])
    assert torch.equal(a.rendering(nxt, ["roughness.value"], channel_first=True),
                       b.rendering(nxt, ["roughness.value"], channel_first=True))
    with pytest.raises(NotImplementedError):
        rendering_refmaps(b, ["name"], z)
