#!/usr/bin/env python3
"""Times clumpy_cuda_age_offsets_dev (the reset offsets of advect_points.cc:127-135 drawn on the device) for the C4
point count and for one eighth of it: CUDA events around the two launches."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, clumpy_b200
ctx = clumpy_b200.Context(0)
out = torch.zeros(16_777_216 + 8, dtype=torch.int32, device="cuda")
for n in (16_777_216, 2_097_152):
    for rep in range(3):
        ctx.timer_start(); ctx.age_offsets_dev(1000, n, out.data_ptr()); print("age_offsets_dev", n, "draws:", round(ctx.timer_stop(), 3), "ms")
