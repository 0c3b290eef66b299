"""Per-pass traversal statistics of one render per footprint class (development tool, GPU)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200 import _lib
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
env = torch.from_numpy(synthetic_envmap(1000, 2000, seed=1001)).cuda()[None]
v = torch.tensor([[0.3, 0.0, 1.0]])
for rough in [float(x) for x in sys.argv[1:]] or [0.7, 0.3, 0.15, 0.09, 0.0]:
    for name, kw in (("default", {}), ("rim off", {"limb_x": 0.0})):
        o = _lib.default_render_options(); o.collect_stats = 1
        for k, val in kw.items(): setattr(o, k, val)
        z = torch.tensor([[0.5, 0.9, 0.5, 0.3, rough, 1.0]])
        render_batch(env, z, v, res=128, footprint_S=None, options=o, check_status=True)
        st = render_batch.last_status
        print(f"rough {rough} [{name}]")
        for p in range(5):
            vis, it, a0, a1 = st[16 + 4 * p: 20 + 4 * p]
            if vis:
                nodes = 16384 * 4 ** p
                print(f"  pass {p}: visits {vis/1e6:7.2f} M  iters {it/1e6:6.2f} M (fill {vis/max(it,1)/32:.2f})  texel recs {a0/1e6:7.2f} M  pyramid recs {a1/1e6:7.2f} M"
                      f"  -> per cell: visits {vis*32/16384:9.0f} pairs0 {a0*32/16384:9.0f} pairs1 {a1*32/16384:9.0f}")
