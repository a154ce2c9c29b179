#!/usr/bin/env python3
"""Prints frame agreement of every accumulator type against the golden reference frames and a few
oracle cases (run on a GPU box).  Used to fill the parity table in DESIGN.md."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clumpy_b200, oracle

G = os.path.join(ROOT, "tests", "golden")
meta = json.load(open(os.path.join(G, "golden.json")))
adv = dict(np.load(os.path.join(G, "advect_small.npz")))

def agree(a, b):
    d = np.abs(a.astype(int) - b.astype(int))
    return float((d == 0).mean()), float((d <= 1).mean()), int(d.max())

ctx = clumpy_b200.Context(0)
rows = []
for name in ["k1", "k3", "k5", "k7", "k3_wrap", "k9"]:
    c = meta[f"advect_{name}"]
    for cells in ["exact", "f16", "u8", "u32"]:
        os.environ["CLUMPY_CUDA_CELLS"] = cells
        fr = ctx.advect_points(adv["pts"], adv[c["vel"]], c["step"], c["k"], c["decay"], c["nframes"])
        worst = min(agree(fr[f], adv[f"frames_{name}"][f]) for f in range(len(fr)))
        mx = max(agree(fr[f], adv[f"frames_{name}"][f])[2] for f in range(len(fr)))
        rows.append((f"golden {name}", cells, worst[0], worst[1], mx))
# example6-like: sparse points, k=5, long trails, pendulum-like converging field
rng = np.random.default_rng(1)
w, h, n, nf = 800, 400, 2500, 60
yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
th = (xx / w - 0.5) * 4 * np.pi; om = (yy / h - 0.5) * 5
vel = np.stack([om, -0.9 * om - np.sin(th)], axis=-1).astype(np.float32)
pts = (rng.random((n, 2)) * [w, h]).astype(np.float32)
ref, _ = oracle.advect_points(pts, vel, 2.5, 5, 0.99, nf)
for cells in ["exact", "f16", "u32"]:
    os.environ["CLUMPY_CUDA_CELLS"] = cells
    fr = ctx.advect_points(pts, vel, 2.5, 5, 0.99, nf)
    ag = [agree(fr[f], ref[f]) for f in range(nf)]
    rows.append(("example6-like 800x400 k5", cells, min(a[0] for a in ag), min(a[1] for a in ag), max(a[2] for a in ag)))
# dense k=3 (C4-like density 0.5 pts/px)
w, h, n, nf = 512, 256, 65536, 30
pot, _, _ = oracle.generate_simplex(w, h, 1.0, 4.0, 0); vel, _, _ = oracle.curl_2d(pot)
pts = (rng.random((n, 2)) * [w, h]).astype(np.float32)
ref, _ = oracle.advect_points(pts, vel, 30.0, 3, 0.99, nf)
for cells in ["f16", "u8", "u32"]:
    os.environ["CLUMPY_CUDA_CELLS"] = cells
    fr = ctx.advect_points(pts, vel, 30.0, 3, 0.99, nf)
    ag = [agree(fr[f], ref[f]) for f in range(nf)]
    rows.append(("C4-like 512x256 k3 0.5pt/px", cells, min(a[0] for a in ag), min(a[1] for a in ag), max(a[2] for a in ag)))
print(f"{'case':34s} {'cells':6s} {'== (worst frame)':>17s} {'<=1 LSB':>9s} {'max':>4s}")
for r in rows:
    print(f"{r[0]:34s} {r[1]:6s} {r[2]:17.5f} {r[3]:9.5f} {r[4]:4d}")
