"""tree vs flat, local relative error over the whole image under option sets (development tool)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200 import _lib
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
g = np.load(sys.argv[1])
for i in [int(x) for x in sys.argv[2:]]:
    He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
    seed, zi, vi, S, nc = [int(x) for x in g["meta"][i]]
    env = torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
    z = torch.tensor(g["z"][i], dtype=torch.float32)[None]; v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
    kw = dict(res=res, footprint_S=S, alpha_min=float(g["alpha_min"]), channel_first=False)
    b = render_batch(env, z, v, flat=True, **kw)[0].double().cpu().numpy()
    print(f"case {i}: seed {seed} z{zi} v{vi} S={S} rough {float(z[0,4]):.3f} peak {b.max():.1f} median {np.median(b):.3f}")
    def run(name, **o_kw):
        o = _lib.default_render_options()
        for k, val in o_kw.items(): setattr(o, k, val)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        render_batch(env, z, v, options=o, **kw)
        e0.record()
        a = render_batch(env, z, v, options=o, check_status=True, **kw)[0].double().cpu().numpy()
        e1.record(); torch.cuda.synchronize()
        loc = np.abs(a - b).max(2) / (np.abs(b).max(2) + 1e-2 * np.median(b))
        k = np.unravel_index(loc.argmax(), loc.shape)
        print(f"  {name:34s} rel-L2 {np.linalg.norm(a-b)/np.linalg.norm(b):.2e} local max {loc.max():.2e} at {k} p99.9 {np.quantile(loc,0.999):.2e}  {e0.elapsed_time(e1):.2f} ms")
    run("default")
    run("limb_hand 8", limb_hand=8.0)
    run("limb_hand 4", limb_hand=4.0)
    run("limb_hand 2", limb_hand=2.0)
    run("limb_hand 1", limb_hand=1.0)
    run("limb_hand 4 boost 3", limb_hand=4.0, limb_boost=3.0)
