"""tree vs flat for one golden case: error by column / row (development tool)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from drmnet_b200.renderer import render_batch
from drmnet_b200.synth import synthetic_envmap
g = np.load(sys.argv[1]); i = int(sys.argv[2])
He, We, res = int(g["He"]), int(g["We"]), int(g["res"])
seed, zi, vi, S, nc = [int(x) for x in g["meta"][i]]
env = torch.from_numpy(synthetic_envmap(He, We, seed=seed)).cuda()[None]
z = torch.tensor(g["z"][i], dtype=torch.float32)[None]; v = torch.tensor(g["view"][i], dtype=torch.float32)[None]
kw = dict(res=res, footprint_S=S, alpha_min=float(g["alpha_min"]), channel_first=False)
a = render_batch(env, z, v, check_status=True, **kw)[0].double().cpu().numpy()
b = render_batch(env, z, v, flat=True, **kw)[0].double().cpu().numpy()
cl = g["cells"][i][:nc]; ref = g["values"][i][:nc]
print("flat vs golden", np.linalg.norm(b[cl[:,0],cl[:,1]]-ref)/np.linalg.norm(ref), "tree vs golden", np.linalg.norm(a[cl[:,0],cl[:,1]]-ref)/np.linalg.norm(ref))
print("tree vs flat whole", np.linalg.norm(a-b)/np.linalg.norm(b))
e = np.abs(a-b).max(2)
print("col max err / peak:", np.round(e.max(0)/np.abs(b).max()*1e4,1).tolist())
print("row max err / peak:", np.round(e.max(1)/np.abs(b).max()*1e4,1).tolist())
idx = np.dstack(np.unravel_index(np.argsort(-e.ravel())[:10], e.shape))[0]
for (r,c) in idx: print(r,c,'tree',a[r,c].round(4),'flat',b[r,c].round(4))
