"""Diffuse lattice correction on/off: error against each other and against the un-accelerated sum (GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drmnet_b200 import synth
from drmnet_b200.renderer import render_batch, auto_footprint

def rel(a, b):
    return (torch.linalg.norm((a - b).flatten(1), dim=1) / torch.linalg.norm(b.flatten(1), dim=1))

def main():
    dev = "cuda:0"
    He, We = 1000, 2000
    B = 8
    e0 = int(sys.argv[1]) if len(sys.argv) > 1 else 100   # envmap seed offset
    z0 = int(sys.argv[2]) if len(sys.argv) > 2 else 500   # BRDF / view seed offset
    R = int(sys.argv[3]) if len(sys.argv) > 3 else B * 3  # renders
    env = torch.stack([synth.synthetic_envmap(He, We, e0 + i, device=dev) for i in range(B)])
    z = torch.stack([synth.sample_brdf(z0 + i) for i in range(R)]).to(dev)
    view = torch.stack([synth.sample_view(z0 + i) for i in range(R)]).to(dev)
    idx = (torch.arange(R, device=dev) % B).int()
    def run():
        torch.cuda.synchronize(); t = time.time()
        o = render_batch(env, z, view, env_index=idx, res=128, footprint_S=None)
        torch.cuda.synchronize()
        return o, time.time() - t
    run()
    on, t_on = run()
    os.environ["DRM_RENDER_DIFF_CORR"] = "0"
    run()
    off, t_off = run()
    for k in ("DRM_RENDER_COARSE", "DRM_RENDER_LEVELS", "DRM_RENDER_NEAR", "DRM_RENDER_FAR_COARSE", "DRM_RENDER_FAR_COARSE4", "DRM_RENDER_UNIFY"):
        os.environ[k] = "0"
    full, _ = run()
    print("t_on %.3f t_off %.3f" % (t_on, t_off))
    r1, r2, r3 = rel(on, off), rel(on, full), rel(off, full)
    print("on vs off  max %.3e" % r1.max().item())
    print("on vs full max %.3e  off vs full max %.3e" % (r2.max().item(), r3.max().item()))
    for i in range(R):
        print(i, "r=%.3f m=%.2f" % (z[i, 4].item(), z[i, 0].item()), "%.2e %.2e %.2e" % (r1[i].item(), r2[i].item(), r3[i].item()))

if __name__ == "__main__":
    main()
