"""Time of ONE render call (N = 1) per footprint class, device events, against the sum of its kernels (development tool)."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200.renderer import B200RefMapRenderer, render_batch
from drmnet_b200.synth import BRDF_PARAM_NAMES, synthetic_envmap
env = torch.from_numpy(synthetic_envmap(1000, 2000, seed=1001)).cuda()
v = torch.tensor([[0.3, 0.0, 1.0]])
for rough in (0.7, 0.3, 0.15):
    z = torch.tensor([[0.5, 0.9, 0.5, 0.3, rough, 1.0]])
    for _ in range(3):
        render_batch(env[None], z, v, res=128, footprint_S=None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        render_batch(env[None], z, v, res=128, footprint_S=None)
    e1.record(); torch.cuda.synchronize()
    print(f"render_batch N=1 rough {rough}: device {e0.elapsed_time(e1) / 20:.3f} ms per call, host wall {(time.perf_counter() - t0) * 50:.3f} ms")
r = B200RefMapRenderer(refmap_res=128, spp=256, envmap_size=(1000, 2000), denoise="simple", brdf_param_names=BRDF_PARAM_NAMES)
r.rendering(torch.tensor([0.5, 0.9, 0.5, 0.3, 0.3, 1.0]), BRDF_PARAM_NAMES, envmap=env, channel_first=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20):
    r.rendering(torch.tensor([0.5, 0.9, 0.5, 0.3, 0.3 + 0.01 * i, 1.0]), BRDF_PARAM_NAMES, channel_first=True)
torch.cuda.synchronize()
print(f"stateful rendering(envmap=None): {(time.perf_counter() - t0) * 50:.3f} ms per call")
