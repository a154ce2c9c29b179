#!/usr/bin/env python3
"""Where the wall clock of one whole `advect_points` command goes (CLUMPY_CUDA_TRACE=1 marks on stderr): the C4
inputs in pinned host memory, three calls of clumpy_cuda_advect_points with 10 recorded frames each."""
import os
import sys
import time

os.environ["CLUMPY_CUDA_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import clumpy_b200  # noqa: E402
from bench import C4, synth_points  # noqa: E402

cfg = dict(C4)
n, w, h = cfg["npts"], cfg["width"], cfg["height"]
ctx = clumpy_b200.Context(0)
vel = torch.zeros((h, w, 2), dtype=torch.float32).pin_memory()
vel += 0.001
pts = torch.from_numpy(synth_points(n, w, h)).pin_memory()
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.advect_points(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], 10, on_frame=lambda f, px: 0)
    torch.cuda.synchronize()
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms", file=sys.stderr)
ctx.close()
