"""One batch of renders of a single footprint class (for ncu captures).  usage: tree_one.py <roughness> [n] [He]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
rough = float(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 8; He = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
envs = torch.stack([torch.from_numpy(synthetic_envmap(He, 2 * He, seed=1000 + b)) for b in range(n)]).cuda()
view = torch.tensor([[0.3, 0.0, 1.0]]).repeat(n, 1)
z = torch.tensor([[0.5, 0.9, 0.5, 0.3, rough, 1.0]]).repeat(n, 1)
for _ in range(2):
    out = render_batch(envs, z, view, res=128, footprint_S=None)
torch.cuda.synchronize()
print(float(out.sum()))
