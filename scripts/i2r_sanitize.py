import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from drmnet_b200.img2refmap import img2refmap_batch
from drmnet_b200.synth import sphere_image_inputs
dev="cuda:0"
c, n = sphere_image_inputs(24, seed=2)
offs = torch.tensor([0, len(c) // 2, len(c)], dtype=torch.int64, device=dev)
for mode in ("median", "mean"):
    r = img2refmap_batch(torch.from_numpy(c).to(dev), torch.from_numpy(n).to(dev), offs, 16, float(np.pi / 32), reduce=mode)
    torch.cuda.synchronize()
    print("img2refmap", mode, int(r[1].sum()))
