#!/usr/bin/env python3
"""A short, steady-state run of the single-GPU advect pipeline on the C4 workload for ncu: device-resident inputs,
`warm` frames (default 1208: past one reset period, so the cloud has clumped into its steady state), then `frames`
more (default 16).  With 8 frames per launch, launch number warm / 8 + 1 of advect_fused_kernel and of
composite_group_kernel is the first one after the warm-up:

    ncu --set full --clock-control none --import-source on -k regex:'advect_fused|composite_group' -s 302 -c 2 \\
        -o gpurun_out/prof python tools/profile_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import clumpy_b200  # noqa: E402
from bench import C4, synth_points  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1208
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = dict(C4)
n, w, h = cfg["npts"], cfg["width"], cfg["height"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = clumpy_b200.Context(0, stream.cuda_stream)
d_pot = torch.empty((h, w), dtype=torch.float32, device="cuda")
d_vel = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
ctx.generate_simplex_dev(w, h, 0, h, cfg["amplitude"], cfg["frequency"], cfg["seed"], d_pot.data_ptr())
ctx.curl_2d_dev(d_pot.data_ptr(), w, h, None, d_vel.data_ptr())
torch.cuda.synchronize()
d_pts = torch.from_numpy(synth_points(n, w, h)).cuda()
adv = ctx.advect(d_pts.data_ptr(), d_vel.data_ptr(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], cfg["nframes"],
                 npts=n, width=w, height=h)
adv.step(warm)
torch.cuda.synchronize()
ctx.timer_start()
adv.step(frames)
ms = ctx.timer_stop()
print(f"{frames} frames after {warm}: {ms / frames:.4f} ms per frame")
adv.close()
ctx.close()
