#!/usr/bin/env python3
"""Multi-rank parity check of the sharded advect path.  Launch on a box with N GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py

Every rank takes its shard of the points; after every frame the gathered framebuffer and the gathered
trajectories are compared with the CPU oracle on rank 0 (frames: +-1 LSB on >= 99.9 % of pixels;
trajectories: bit-identical).  Exits non-zero on any mismatch."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clumpy_b200  # noqa: E402
import oracle  # noqa: E402
from clumpy_b200.multigpu import make_sharded  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = clumpy_b200.Context(local, stream.cuda_stream)
    ok = True
    # generate_simplex -> curl_2d with the rows sharded across the ranks (+ all-gather of the velocity
    # bands) must reproduce the single-GPU field bit for bit
    from clumpy_b200.multigpu import sharded_simplex_curl
    for (w, h, freq, seed) in [(512, 64, 8.0, 0), (300, 37, 20.0, 5), (2048, 1024, 8.0, 1)]:
        got = sharded_simplex_curl(ctx, w, h, 1.0, freq, seed, dist, device).cpu().numpy()
        pot = torch.empty((h, w), dtype=torch.float32, device=device)
        whole = torch.empty((h, w, 2), dtype=torch.float32, device=device)
        ctx.generate_simplex_dev(w, h, 0, h, 1.0, freq, seed, pot.data_ptr())
        ctx.curl_2d_dev(pot.data_ptr(), w, h, None, whole.data_ptr())
        torch.cuda.synchronize()
        same = got.shape == (h, w, 2) and np.array_equal(got.view(np.uint32), whole.cpu().numpy().view(np.uint32))
        ok = ok and same
        if rank == 0:
            print(f"[fields] simplex+curl {w}x{h} row-sharded x{dist.get_world_size()}: identical to one GPU = {same}")
    kinds = os.environ.get("CLUMPY_MULTIGPU_KINDS", "routed,dense").split(",")
    for kind in kinds:
      for (w, h, n, k, decay, nframes, step) in [(320, 200, 70001, 3, 0.98, 6, 35.0), (256, 160, 30000, 1, 0.9, 5, 40.0),
                                               (200, 136, 20000, 5, 0.95, 4, 30.0), (512, 64, 50000, 3, -0.97, 4, 60.0)]:
          pot, _, _ = oracle.generate_simplex(w, h, 1.0, 5.0, 8)
          vel, _, _ = oracle.curl_2d(pot)
          pts = ((np.random.default_rng(21).random((n, 2)) * 1.04 - 0.02) * [w, h]).astype(np.float32)
          run = make_sharded(kind, ctx, pts, vel, step, k, decay, nframes, dist, device)
          ref = oracle.Advect(pts, vel, step, k, decay, nframes) if rank == 0 else None
          s = 0
          for chunk in (1, 8, 3, 8, 8, 8):  # groups of frames: one route launch + one drain + one composite each
              chunk = min(chunk, 2 * nframes - s)
              if chunk == 0:
                  break
              run.step(chunk)
              s += chunk
              img = run.gather_image()
              traj = run.points()
              if rank == 0:
                  for _ in range(chunk):
                      ref.step()
                  d = np.abs(img.astype(int) - ref.image().astype(int))
                  frac = float((d <= 1).mean())
                  same = np.array_equal(traj.view(np.uint32), ref.points().view(np.uint32))
                  if frac < 0.999 or not same:
                      ok = False
                      print(f"MISMATCH [{kind}] {w}x{h} k={k} frame {s}: within1={frac:.5f} max={d.max()} traj_equal={same}")
          if rank == 0:
              print(f"[{kind}] case {w}x{h} n={n} k={k} decay={decay}: ok={ok} owned rows per rank: "
                    f"{(run.y0, run.y1)}")
          run.close()
    flag = torch.tensor([0 if ok else 1], device=device)
    dist.all_reduce(flag)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
