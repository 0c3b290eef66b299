"""CPU (numpy, fp64) study of the per-cell level scheme at limb cells: the pixel sum split into the distance bands of
render_near_kernel, each band evaluated on its own lattice (with the view-term mean and centroid shift of
view_term_avg: mode "shift"; mean only: "avg"; neither: "plain") against the full 16x16 lattice.

usage: limb_levels_study.py ALPHA [shift|avg|plain] [THRESHOLD_FACTOR] [limb] [MIN_LATTICE]   (500x1000 synthetic map, res 128;
       minutes on one core; "limb": only the last four columns of row 64)
"""
import math, sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from drmnet_b200.synth import synthetic_envmap
He, We, res, S = 500, 1000, 128, 16
alpha = float(sys.argv[1]) if len(sys.argv) > 1 else 0.09
mode = sys.argv[2] if len(sys.argv) > 2 else "shift"
env = synthetic_envmap(He, We, 1004).astype(np.float64).sum(-1)
t = (np.arange(He) + 0.5) * math.pi / He; p = (np.arange(We) + 0.5) * 2 * math.pi / We
st, ct = np.sin(t)[:, None], np.cos(t)[:, None]
d = np.stack([st * np.sin(p)[None], np.broadcast_to(ct, (He, We)), -st * np.cos(p)[None]], -1).reshape(-1, 3)
dom = (math.pi / He) * (2 * math.pi / We) * np.broadcast_to(st, (He, We)).reshape(-1)
E = env.reshape(-1) * dom
v = np.array([0.0, 0.0, 1.0]); left = np.array([1.0, 0, 0]); up = np.array([0, 1.0, 0])
c = math.pi / res
hvec = v[None] + d; ln = np.linalg.norm(hvec, axis=1); h = hvec / np.maximum(ln, 1e-12)[:, None]
a2 = alpha * alpha
GL = {s: (np.polynomial.legendre.leggauss(s)[0], np.polynomial.legendre.leggauss(s)[1] / 2) for s in (1, 2, 4, 8, 16)}
def nrm(th, ph): return math.sin(th) * math.cos(ph) * left + math.cos(th) * up + math.sin(th) * math.sin(ph) * v
def V(lz): return 1.0 / (math.pi * a2 * (lz + math.sqrt(lz * lz * (1 - a2) + a2)))
def T(n, mask):
    nh = h[mask] @ n; xx = ln[mask] * nh - n @ v; xc = np.maximum(xx, 0)
    qq = 1 + (1 - nh * nh) * (1 / a2 - 1)
    return float(np.sum(xc / (qq * qq * (xc + np.sqrt(xc * xc * (1 - a2) + a2))) * E[mask]))
ls, ns = 0.3, 0.6
thr = [21 * math.sqrt(c * alpha), 7.5 * c ** (2 / 3) * alpha ** (1 / 3), 2.4 * c ** 0.8 * alpha ** 0.2, 1.2 * c ** (8 / 9) * alpha ** (1 / 9)]
thr = [ls * max(x, 6 * alpha) for x in thr]
fac = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
tk = [fac * x * (ns / ls) + 0.75 * c for x in thr] + [0.0]
xf, wf = GL[S]
MINLAT = int(sys.argv[5]) if len(sys.argv) > 5 else 1
def node(i, j, Sk, a, b):
    m = S // Sk
    xs, ws = GL[Sk]
    if m == 1 or mode == "plain":
        th, ph = (i + .5 + .5 * xs[a]) * c, (j + .5 + .5 * xs[b]) * c
        return nrm(th, ph), ws[a] * ws[b] * V(math.sin(th) * math.sin(ph))
    num = den = va = vb = ua = ub = 0.0
    for ia in range(m):
        for ib in range(m):
            xa, xb = xf[a * m + ia], xf[b * m + ib]; w = wf[a * m + ia] * wf[b * m + ib]
            wv = w * V(math.sin((i + .5 + .5 * xa) * c) * math.sin((j + .5 + .5 * xb) * c))
            num += wv; den += w; va += wv * xa; vb += wv * xb; ua += w * xa; ub += w * xb
    sa, sb = (va / num - ua / den, vb / num - ub / den) if mode == "shift" else (0.0, 0.0)
    return nrm((i + .5 + .5 * (xs[a] + sa)) * c, (j + .5 + .5 * (xs[b] + sb)) * c), ws[a] * ws[b] * num / den
CELLS = ((64, 127), (64, 126), (64, 125), (64, 124)) if len(sys.argv) > 4 else ((64, 126), (64, 127), (63, 125), (64, 64), (20, 100))
for (i, j) in CELLS:
    n1, _ = node(i, j, 1, 0, 0)
    chord = np.linalg.norm(h - n1[None], axis=1)
    vis = ln > 0.05
    exact_all = 0.0; per = []
    bands = []
    for k, Sk in enumerate((1, 2, 4, 8, 16)):
        hi = np.inf if k == 0 else tk[k - 1]
        lo = tk[k]
        bands.append(vis & (chord >= lo) & (chord < hi))
    ex = np.zeros(5); ap = np.zeros(5)
    for a in range(S):
        for b in range(S):
            n, wv = node(i, j, S, a, b)
            for k in range(5):
                if bands[k].any(): ex[k] += wv * T(n, bands[k])
    for k, Sk in enumerate((1, 2, 4, 8, 16)):
        if not bands[k].any(): continue
        Sk = max(Sk, MINLAT)
        for a in range(Sk):
            for b in range(Sk):
                n, wv = node(i, j, Sk, a, b)
                ap[k] += wv * T(n, bands[k])
    tot = ex.sum()
    print(f"cell {i},{j} n.v={n1 @ v:.4f} band shares {np.round(ex / tot, 3)}  err/total per band {np.array2string((ap - ex) / tot, precision=2)}  sum {((ap - ex).sum()) / tot:+.2e}")
