import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, clumpy_b200
ctx = clumpy_b200.Context(0)
sw, sh = 16384, 8192
rng = np.random.default_rng(0)
n = 1 << 20
pts = np.stack([rng.random(n, dtype=np.float32) * sw, rng.random(n, dtype=np.float32) * sh], axis=1)
for mode in ("auto", "u32", "auto", "u32"):
    if mode == "auto": os.environ.pop("CLUMPY_CUDA_CELLS", None)
    else: os.environ["CLUMPY_CUDA_CELLS"] = mode
    for rep in range(2):
        ctx.sync(); t0 = time.perf_counter(); img = ctx.splat_disks(pts, sw, sh, 1.0, 3); ctx.sync()
        print(f"cells={mode} rep {rep}: {time.perf_counter() - t0:.4f} s nonzero {np.count_nonzero(img)}", flush=True)
ctx.close()
