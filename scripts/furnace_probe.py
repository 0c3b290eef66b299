"""White-furnace deviation of the canonical quadrature at full size, by footprint S (diagnostic, GPU)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import Z0
dev = "cuda:0"
env = torch.ones(1, 1000, 2000, 3, device=dev)
v = torch.tensor([[0.0, 0.0, 1.0]])
for am in (0.0, 0.004, 0.008):
    for S in (1, 2, 4, 8, 16):
        torch.cuda.synchronize(); t = time.time()
        w = render_batch(env, torch.tensor([list(Z0)]), v, res=128, footprint_S=S, alpha_min=am)[0]
        torch.cuda.synchronize(); dt = time.time() - t
        dev_ = (w - 1).abs()
        print(f"alpha_min={am} S={S:2d} time={dt*1e3:8.1f} ms  max|w-1| margin 8: {dev_[:, 8:-8, 8:-8].max():.4f}  16: {dev_[:, 16:-16, 16:-16].max():.4f}  32: {dev_[:, 32:-32, 32:-32].max():.4f}  rms32: {dev_[:, 32:-32, 32:-32].pow(2).mean().sqrt():.5f}")
