#!/usr/bin/env python3
"""Throughput of the kernels around advect_points on ONE B200 (Mpx/s for generate_simplex and
curl_2d at the C3 size, Mpts/s for the splat_points sweep C5).  One JSON object per line on stdout.

    python tools/bench_aux.py [--quick]

simplex / curl: device-resident (clumpy_cuda_generate_simplex_dev / _curl_2d_dev), CUDA events on the
launching stream, 16384 x 16384, mean of 5 launches after 2 warm-ups.  curl reads 4 and writes 8 bytes per
pixel (12 B/px algorithmic) against the measured HBM copy bandwidth; simplex writes 4 B/px and is FP32/INT
issue bound.  splat sweep: the whole clumpy_cuda_splat_disks / _splat_points call with HOST buffers (H2D of
the points, the 16384 x 8192 u8 image back), wall clock with the device idle on both sides -- the number a
`clumpy splat_points` user sees minus file I/O."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device memory only)

import clumpy_b200  # noqa: E402
from bench import measured_peaks  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    hbm, src = measured_peaks()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = clumpy_b200.Context(0, stream.cuda_stream)
    w = h = 4096 if quick else 16384
    pot = torch.empty((h, w), dtype=torch.float32, device="cuda")
    vel = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    for name, fn, bytes_px in (
        ("generate_simplex", lambda: ctx.generate_simplex_dev(w, h, 0, h, 1.0, 8.0, 0, pot.data_ptr()), 4),
        ("curl_2d", lambda: ctx.curl_2d_dev(pot.data_ptr(), w, h, None, vel.data_ptr()), 12),
    ):
        for _ in range(2):
            fn()
        ctx.timer_start()
        for _ in range(5):
            fn()
        ms = ctx.timer_stop() / 5
        gbs = w * h * bytes_px / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "size": f"{w}x{h}", "ms": ms, "Mpx_per_s": w * h / (ms * 1e-3) / 1e6,
                          "algorithmic_GB_per_s": gbs, "frac_of_hbm": gbs / hbm, "hbm_peak_GB_per_s": hbm,
                          "peak_source": src}), flush=True)
    del pot, vel
    sw, sh = (4096, 2048) if quick else (16384, 8192)
    rng = np.random.default_rng(0)
    for logn in ((20,) if quick else (20, 22, 24, 26)):
        n = 1 << logn
        x = rng.random(n, dtype=np.float32) * np.float32(sw)
        y = rng.random(n, dtype=np.float32) * np.float32(sh)
        pts = torch.from_numpy(np.stack([x, y], axis=1)).pin_memory()
        for k in (1, 3, 7):
            def call():
                if k == 1:  # `splat_points ... u8disk 1 1.0` takes the fast path (splat_points.cc:93-94)
                    return ctx.splat_points(pts.numpy(), sw, sh)
                return ctx.splat_disks(pts.numpy(), sw, sh, 1.0, k)
            call()
            dt = float("inf")
            for _ in range(3):  # best of 3: a cudaMalloc of a GB-sized counter image now and then takes 0.5 s
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                img = call()
                torch.cuda.synchronize()
                dt = min(dt, time.perf_counter() - t0)
            print(json.dumps({"kernel": "splat_points (host buffers)", "image": f"{sw}x{sh}", "npts": n, "kernel_size": k,
                              "seconds": dt, "Mpts_per_s": n / dt / 1e6, "nonzero_px": int(np.count_nonzero(img)),
                              "max": int(img.max())}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
