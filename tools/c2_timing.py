import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, clumpy_b200
ctx = clumpy_b200.Context(0)
w, h = 4000, 2000
rng = np.random.default_rng(0)
pts = (rng.random((12392, 2)) * [w, h]).astype(np.float32)
vel = (rng.standard_normal((h, w, 2)) * 0.1).astype(np.float32)
adv = ctx.advect(pts, vel, 2.5, 5, 0.99, 400)
adv.step(5); ctx.sync()
t0 = time.perf_counter(); adv.step(40); ctx.sync(); dt = time.perf_counter() - t0
print(f"C2-like advect (12392 pts, 4000x2000, k=5): {dt/40*1e3:.3f} ms per frame; profile advect/composite/clear ms: {adv.profile(10)}")
adv.close(); ctx.close()
