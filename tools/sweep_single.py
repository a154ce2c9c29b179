#!/usr/bin/env python3
"""A/B sweep of the single-GPU advect pipeline on the C4 workload: frames per launch, composite overlap, composite
CTA size, sort tile shape.  Each configuration: 1200 warm-up frames (one reset period: the cloud has clumped into
its steady state), then 640 timed frames (ten regroup periods), CUDA events.  One JSON object per line.

    python tools/sweep_single.py [quick]
"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import clumpy_b200  # noqa: E402
from bench import C4, synth_points  # noqa: E402


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
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
    d_off = torch.from_numpy(clumpy_b200.age_offsets(cfg["nframes"], n).astype("int32")).cuda()
    # (round 2's sweep b also had an "8 CTAs of 32 registers per SM" variant of the advect kernel -- the advect_minb
    # column of profiles/r2_sweep_single_b.jsonl; it lost and the switch is gone)
    base = dict(fuse=8, overlap=-1, threads="512", tile="4,1", hints="0", regroup=0)
    variants = [{}] + [dict(hints=str(hb)) for hb in (1, 2, 3, 4, 8, 12, 5, 15)] + [dict(tile="3,2"), dict(regroup=96), dict(regroup=128),
                dict(fuse=6), dict(overlap=1)]
    if quick:
        variants = variants[:3]
    for var in variants:
        c = dict(base, **var)
        fuse, overlap, thr, tile = c["fuse"], c["overlap"], c["threads"], c["tile"]
        os.environ["CLUMPY_CUDA_COMPOSITE_THREADS"] = thr
        os.environ["CLUMPY_CUDA_SORT_TILE"] = tile
        # (the library reads these two once per process: they are fixed per run of this script unless re-exec'd)
        os.environ["CLUMPY_CUDA_HINTS"] = c["hints"]
        adv = ctx.advect(d_pts.data_ptr(), d_vel.data_ptr(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], cfg["nframes"],
                         age_offset=d_off.data_ptr(), npts=n, width=w, height=h, fuse=fuse, overlap=overlap, regroup=c["regroup"])
        adv.step(1200)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        adv.step(640)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 640
        a_live, c_live, nf = adv.profile_stages(16)
        a_alone, c_alone, _ = adv.profile_stages(16, serial=True)
        adv.close()
        print(json.dumps({"fuse": fuse, "overlap": overlap, "composite_threads": int(thr), "sort_tile_log2": tile, "hints": c["hints"],
                          "regroup": c["regroup"], "ms_per_step": ms,
                          "advect_ms_per_frame_live": a_live / nf, "composite_ms_per_frame_live": c_live / nf,
                          "advect_ms_per_frame_alone": a_alone / nf, "composite_ms_per_frame_alone": c_alone / nf}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
