"""One golden case under several option sets (development tool): tree_case.py <npz> <case index>"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200 import _lib
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
g = np.load(sys.argv[1]); 
for i in [int(x) for x in sys.argv[2:]]:
    He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
    seed, zi, vi, S, nc = [int(x) for x in g["meta"][i]]
    env = torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
    z = torch.tensor(g["z"][i], dtype=torch.float32)[None]; v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
    cl = g["cells"][i][:nc]; ref = g["values"][i][:nc]
    print(f"case {i}: seed {seed} z{zi} v{vi} S={S} rough {float(z[0,4]):.3f}")
    def run(name, flat=False, **kw):
        o = _lib.default_render_options()
        for k, val in kw.items(): setattr(o, k, val)
        out = render_batch(env, z, v, res=res, footprint_S=S, alpha_min=float(g["alpha_min"]), channel_first=False,
                           options=o, flat=flat, check_status=not flat)[0].double().cpu().numpy()
        got = out[cl[:, 0], cl[:, 1]]
        e = np.abs(got - ref).max(1); k = int(e.argmax())
        print(f"  {name:28s} rel-L2 {np.linalg.norm(got-ref)/np.linalg.norm(ref):.2e} max/peak {e.max()/np.abs(ref).max():.2e} at cell {cl[k]} marks {getattr(render_batch,'last_status',None)}")
    run("default")
    run("flat", flat=True)
    run("kappa .07", kappa=0.07)
    run("alpha_full2 .1", alpha_full2=0.1)
    run("flat_scale 2.5", flat_scale=2.5)
    run("no pixcov", pixel_covariance=0)
    run("ls 1.2", level_scale=1.2, level_scale0=1.2)
    run("hz .015", horizon=0.015)
    run("rcap .03", rcap=0.03)
