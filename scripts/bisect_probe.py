"""Which acceleration stage carries the error of a given render?  Toggles each DRM_RENDER_* switch alone (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drmnet_b200 import synth
from drmnet_b200.renderer import render_batch, auto_footprint

SW = ["DRM_RENDER_COARSE", "DRM_RENDER_LEVELS", "DRM_RENDER_NEAR", "DRM_RENDER_FAR_COARSE", "DRM_RENDER_DIFF_CORR", "DRM_RENDER_VIEW_AVG", "DRM_RENDER_FAR_COARSE4"]

def rel(a, b):
    return (torch.linalg.norm((a - b).flatten(1), dim=1) / torch.linalg.norm(b.flatten(1), dim=1))

def main():
    dev = "cuda:0"
    He, We = 1000, 2000
    B = 8
    picks = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,5,13,21").split(",")]
    env = torch.stack([synth.synthetic_envmap(He, We, 100 + i, device=dev) for i in range(B)])
    z = torch.stack([synth.sample_brdf(500 + i) for i in picks]).to(dev)
    view = torch.stack([synth.sample_view(500 + i) for i in picks]).to(dev)
    idx = torch.tensor([i % B for i in picks], device=dev).int()
    S = [auto_footprint(float(z[i, 4]), 128) for i in range(len(picks))]
    print("picks", picks, "S", S)
    def run(off):
        for k in SW:
            os.environ.pop(k, None)
        for k in off:
            os.environ[k] = "0"
        o = render_batch(env, z, view, env_index=idx, res=128, footprint_S=None)
        torch.cuda.synchronize()
        return o
    full = run(SW)
    base = run([])
    print("all on  :", ["%.2e" % v for v in rel(base, full).tolist()])
    for k in SW:
        o = run([k])
        print("off %-22s:" % k[11:], ["%.2e" % v for v in rel(o, full).tolist()])
    for k in SW:
        o = run([s for s in SW if s != k])
        print("only %-21s:" % k[11:], ["%.2e" % v for v in rel(o, full).tolist()])
    # where is the error?
    d = (base - full).abs().sum(1)
    for i in range(len(picks)):
        m = d[i].argmax().item()
        print(picks[i], "max abs diff at cell", divmod(m, 128), "diff %.3e value %.3e  image max %.3e" %
              (d[i].flatten()[m].item(), full[i].sum(0).flatten()[m].item(), full[i].sum(0).max().item()))

if __name__ == "__main__":
    main()
