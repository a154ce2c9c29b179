#!/usr/bin/env python3
"""Runs the routed (multi-GPU) advect path with a process group of ONE rank on one GPU, on 1/SHARD of
the C4 points over the full image -- what each rank of an N-GPU run executes, minus the peers.  For
profiling the per-rank kernels under ncu (a multi-rank run must not be profiled)."""
import os, sys, socket, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clumpy_b200
from clumpy_b200.multigpu import RoutedAdvect
from bench import C4, synth_points

shard = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
with socket.socket() as s:
    s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = clumpy_b200.Context(0, stream.cuda_stream)
cfg = dict(C4); n, w, h = cfg["npts"], cfg["width"], cfg["height"]
d_pot = torch.empty((h, w), dtype=torch.float32, device=dev); d_vel = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
ctx.generate_simplex_dev(w, h, 0, h, 1.0, 8.0, 0, d_pot.data_ptr()); ctx.curl_2d_dev(d_pot.data_ptr(), w, h, None, d_vel.data_ptr())
vel = d_vel.cpu().numpy(); del d_pot, d_vel
pts = synth_points(n, w, h)
# what rank 0 of a `shard`-rank run simulates: the points whose home row lies in the top 1/shard of the
# image (home-rows sharding), or with a third argument "index" the first n/shard points (all over the image)
if os.environ.get("CLUMPY_ROUTE_LOOPBACK") != "1":  # (in loopback mode RoutedAdvect takes its rank's shard itself)
    pts = pts[: n // shard] if (len(sys.argv) > 3 and sys.argv[3] == "index") else pts[pts[:, 1] < h / shard]
if os.environ.get("ROUTED_SINGLE_PLAIN") == "1":  # the single-GPU kernels on the same subset, for comparison
    adv = ctx.advect(pts, vel, cfg["step_size"], cfg["kernel_size"], cfg["decay"], cfg["nframes"], cells="f16", overlap=-1)
    adv.step(80); torch.cuda.synchronize()
    print(f"plain advect, {len(pts)} points: advect / composite ms per launch, frames per launch: {adv.profile_stages(12, serial=True)}")
    adv.close(); ctx.close(); dist.destroy_process_group(); sys.exit(0)
run = RoutedAdvect(ctx, pts, vel, cfg["step_size"], cfg["kernel_size"], cfg["decay"], cfg["nframes"], dist, dev)
run.step(int(os.environ.get("ROUTED_SINGLE_WARM", "80"))); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run.step(steps); e1.record(stream); torch.cuda.synchronize()
print(f"routed single rank, {len(pts)} points on {w}x{h}: {e0.elapsed_time(e1)/steps:.4f} ms/frame; "
      f"route / drain / composite ms per group, frames per group: {run.adv.route_profile(12)}")
run.close(); ctx.close(); dist.destroy_process_group()
