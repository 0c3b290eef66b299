"""Small renders through every launch route for compute-sanitizer runs (memcheck / racecheck), GPU."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import Z0, synthetic_envmap
env = synthetic_envmap(250, 500, 3, device="cuda:0", as_numpy=False)[None]
for S, z in ((16, list(Z0)), (8, [1.0, 0.9, 0.8, 0.7, 0.1, 1.0]), (4, [0.5, 0.9, 0.8, 0.7, 0.15, 1.0]),
             (2, [0.3, 0.8, 0.6, 0.4, 0.3, 0.7]), (1, [0.3, 0.8, 0.6, 0.4, 0.9, 0.7]), (3, [0.3, 0.8, 0.6, 0.4, 0.5, 0.7])):
    o = render_batch(env, torch.tensor([z]), torch.tensor([[0.3, 0.0, 1.0]]), res=24, footprint_S=S)
    torch.cuda.synchronize()
    print(S, float(o.mean()))
