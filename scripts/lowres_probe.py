import os, sys
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap, Z0
dev="cuda:0"
env = synthetic_envmap(250, 500, 1004, device=dev, as_numpy=False)[None]
v = torch.tensor([[0.6442, 0.0, 0.7648]])
def run(res,S,levels,scale):
    os.environ["DRM_RENDER_LEVELS"]="1" if levels else "0"; os.environ["DRM_RENDER_LEVEL_SCALE"]=str(scale)
    return render_batch(env, torch.tensor([list(Z0)]), v, res=res, footprint_S=S)
for res,S in ((24,8),(40,16),(64,16),(128,16)):
    ref=run(res,S,False,1)
    line=f"res {res} S {S}:"
    for sc in (0.3,1.0,3.0,10.0):
        o=run(res,S,True,sc); err=((o-ref).norm()/ref.norm()).item()
        # where is the error
        e=(o-ref).abs().amax(1)[0]; i=int(e.argmax()); 
        line+=f" scale {sc}: {err:.2e} (max at {i//res},{i%res})"
    print(line)
