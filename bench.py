#!/usr/bin/env python3
"""bench.py -- advect_points throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], "C4"): 16 777 216 uniform synthetic points
(numpy default_rng(0)) on an 8192 x 4096 curl-of-simplex velocity field
(generate_simplex 8192x4096 1.0 8.0 0 -> curl_2d), kernel_size 3, step_size 240, decay 0.99,
nframes 1000.  A "step" is ONE simulation frame of advect_points over all points: Euler step +
reset + splat + u8 decay/composite of the 8192 x 4096 framebuffer.

value   point-steps/s with everything resident in HBM, CUDA events on the launching stream.
e2e     the same metric through the public C-ABI call clumpy_cuda_advect_points with HOST (pinned)
        buffers: upload of points + field, K simulation frames, and a device->host copy of each of
        the K/2 recorded 32 MiB frames, wall clock with a device synchronise on both sides.
roofline  HBM roofline of the dominant kernel (advect_kernel: 28 algorithmic bytes per point-step),
        timed live with CUDA events around that kernel alone.
cpu_baseline  the reference CLI (oracle/_ref/clumpy_ref, single-threaded like the reference) on the
        same inputs for a bounded number of frames, on this box's host cores.

--impl reference times only that CPU arm and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

# The advect pipeline keeps kernels on three streams in flight (route / drain + composite / frame copy)
# next to NCCL's and torch's own: give them hardware queues of their own (default 8 per process).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "advect_points point-steps/s"
UNIT = "point-steps/s"
C4 = dict(npts=1 << 24, width=8192, height=4096, kernel_size=3, step_size=240.0, decay=0.99,
          nframes=1000, amplitude=1.0, frequency=8.0, seed=0)
WORKLOAD = ("C4: advect_points 16777216 uniform points (default_rng(0)) on 8192x4096 curl(simplex f=8 seed 0), "
            "kernel 3, step 240, decay 0.99, nframes 1000")
ALGO_BYTES_PER_POINT_STEP = 28  # 8 R + 8 W position, 8 R velocity texel, 4 R reset offset
ALGO_BYTES_PER_PIXEL_FRAME = 2  # u8 framebuffer read + write
# dram__bytes_read.sum + dram__bytes_write.sum of ONE advect_kernel launch on this workload, from the
# round-1 `ncu --set full` capture (profiles/r1_ncu_full_summary.txt: 505.756 MB read + 159.334 MB
# written).  Not measured by this script: a number taken under a profiler is never a bench value.
NCU_TRAFFIC_BYTES_ADVECT_KERNEL_C4 = 505_756_000 + 159_334_000
# the 4-frames-per-launch kernel (profiles/r1_ncu_fused_kernel_summary.txt: 713.551 MB read + 352.322 MB written
# per launch, i.e. per 4 frames of all 16.7 M points: 57 % of its 1879 MB of algorithmic bytes)
NCU_TRAFFIC_BYTES_FUSED_KERNEL_C4 = 713_551_360 + 352_322_304


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_points(npts, width, height):
    """SURVEY 8(d) recipe: x = random(N) * W, then y = random(N) * H, float32, default_rng(0)."""
    rng = np.random.default_rng(0)
    x = rng.random(npts, dtype=np.float32) * np.float32(width)
    y = rng.random(npts, dtype=np.float32) * np.float32(height)
    return np.stack([x, y], axis=1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background during the timed regions."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        # samples under load: the upper half of the observed clocks (idle samples sit at the floor)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference CLI (or, where it could not be built, the oracle port)
# ---------------------------------------------------------------------------------------------

def cpu_reference_run(pts, vel, cfg, nframes_cpu, workdir=None):
    """Runs `advect_points` of the compiled reference for nframes_cpu recorded frames
    (2*nframes_cpu simulation frames) over ALL points; returns (point-steps/s, kind, description)."""
    import oracle
    n = len(pts)
    steps = 2 * nframes_cpu
    if oracle.have_ref():
        base = workdir or ("/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None)
        tmp = tempfile.mkdtemp(prefix="clumpy_ref_", dir=base)
        try:
            np.save(os.path.join(tmp, "pts.npy"), pts)
            np.save(os.path.join(tmp, "vel.npy"), vel)
            # untimed warm-up: pages the binary and the input files in
            oracle.ref_cli("advect_points", "pts.npy", "vel.npy", cfg["step_size"], cfg["kernel_size"],
                           cfg["decay"], 1, "w.npy", cwd=tmp) if n <= (1 << 20) else None
            t0 = time.perf_counter()
            oracle.ref_cli("advect_points", "pts.npy", "vel.npy", cfg["step_size"], cfg["kernel_size"],
                           cfg["decay"], nframes_cpu, "f.npy", cwd=tmp)
            dt = time.perf_counter() - t0
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        kind = "reference"
        what = (f"oracle/_ref/clumpy_ref advect_points (unmodified reference, -O3, 1 thread), whole command incl. "
                f".npy load/save on tmpfs, all {n} points, nframes={nframes_cpu} ({steps} simulation frames)")
    else:
        oracle.lib()
        a = oracle.Advect(pts, vel, cfg["step_size"], cfg["kernel_size"], cfg["decay"], nframes_cpu)
        t0 = time.perf_counter()
        for _ in range(steps):
            a.step()
        dt = time.perf_counter() - t0
        a.close()
        kind = "port"
        what = (f"oracle/liboracle.so restatement (the reference could not be compiled here), 1 thread, compute only, "
                f"all {n} points, {steps} simulation frames")
    return n * steps / dt, kind, what


def host_inputs_for_reference(cfg):
    """C4 inputs built on the CPU (reference CLI when present, oracle otherwise)."""
    import oracle
    pts = synth_points(cfg["npts"], cfg["width"], cfg["height"])
    if oracle.have_ref():
        tmp = tempfile.mkdtemp(prefix="clumpy_field_")
        try:
            oracle.ref_cli("generate_simplex", f"{cfg['width']}x{cfg['height']}", cfg["amplitude"],
                           cfg["frequency"], cfg["seed"], "pot.npy", cwd=tmp)
            oracle.ref_cli("curl_2d", "pot.npy", "vel.npy", cwd=tmp)
            vel = np.load(os.path.join(tmp, "vel.npy"))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    else:
        pot, _, _ = oracle.generate_simplex(cfg["width"], cfg["height"], cfg["amplitude"], cfg["frequency"], cfg["seed"])
        vel, _, _ = oracle.curl_2d(pot)
    return pts, vel


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = dict(C4)
    nframes_cpu = max(1, min(3, args.steps // 2))
    pts, vel = host_inputs_for_reference(cfg)
    value, kind, what = cpu_reference_run(pts, vel, cfg, nframes_cpu)
    steps = 2 * nframes_cpu
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 0, "ms_per_step": 1e3 * cfg["npts"] / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpu_frames": steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import clumpy_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import clumpy_b200.multigpu as mg
        return mg.bench_multi(args, C4, WORKLOAD, METRIC, UNIT)

    cfg = dict(C4)
    n, w, h = cfg["npts"], cfg["width"], cfg["height"]
    K, W = args.steps, args.warmup
    hbm_peak, peak_src = measured_peaks()
    # an explicit (non-null) stream: the library launches on it and torch's events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = clumpy_b200.Context(local_rank, stream.cuda_stream)

    # inputs: field from our own simplex + curl kernels, points from the synthetic recipe (pinned)
    vel = torch.empty((h, w, 2), dtype=torch.float32).pin_memory()
    d_pot = torch.empty((h, w), dtype=torch.float32, device="cuda")
    d_vel = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    ctx.generate_simplex_dev(w, h, 0, h, cfg["amplitude"], cfg["frequency"], cfg["seed"], d_pot.data_ptr())
    ctx.curl_2d_dev(d_pot.data_ptr(), w, h, None, d_vel.data_ptr())
    vel.copy_(d_vel)
    torch.cuda.synchronize()
    del d_pot, d_vel
    pts = torch.from_numpy(synth_points(n, w, h)).pin_memory()
    offsets = torch.from_numpy(clumpy_b200.age_offsets(cfg["nframes"], n)).pin_memory()

    # ---- value: resident in HBM, CUDA events on the launching stream ---------------------------
    adv = ctx.advect(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"],
                     cfg["nframes"], age_offset=offsets.numpy())
    sampler = ClockSampler(local_rank)
    adv.step(W)
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    adv.step(K)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    value = n * K / (ms * 1e-3)

    # ---- roofline: the dominant kernel alone (advect + count), same state, CUDA events --------
    # The dominant kernel as the pipeline runs it: one launch advances `nf` frames (positions stay in
    # registers), timed with CUDA events around the launch while the composites of the previous group
    # share the GPU.  Then the one-frame kernels one by one, for reference.
    launch_ms, nf = adv.profile_fused(40)
    kms, cms, zms = adv.profile(50)
    algo_bytes = n * ALGO_BYTES_PER_POINT_STEP * nf
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    adv.close()

    # ---- e2e: whole command through the C ABI with host buffers ---------------------------------
    nrec = max(1, K // 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.advect_points(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], nrec,
                      on_frame=lambda f, px: 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e_steps = 2 * nrec
    e2e_value = n * e2e_steps / dt
    h2d = (n * 8 + w * h * 8 + n * 4) / e2e_steps
    d2h = (nrec * w * h) / e2e_steps
    clocks = sampler.stop()

    # ---- CPU baseline: bounded sample on this box's host cores ----------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        v, kind, what = cpu_reference_run(pts.numpy(), vel.numpy(), cfg, nframes_cpu=2)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": what,
               "host_cpus": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32+u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (positions 134 MB + field 268 MB + "
                   "counters 134 MB touched every step vs 126 MB L2)", "frames_per_s": K / (ms * 1e-3),
                   "parallelism": "1 GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "seconds": dt,
                "what": "clumpy_cuda_advect_points with pinned host buffers: H2D of points+field+offsets, sort, "
                        f"{e2e_steps} simulation frames, D2H of {nrec} recorded frames"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"advect_fused_kernel ({nf} frames per launch)" if nf > 1 else "advect_kernel",
                     "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_TRAFFIC_BYTES_FUSED_KERNEL_C4 if nf == 4 else (NCU_TRAFFIC_BYTES_ADVECT_KERNEL_C4 if nf == 1 else None),
                     "traffic_source": "ncu --set full capture of this kernel on this workload, per launch (profiles/)",
                     "peak_source": peak_src, "launch_ms": launch_ms, "frames_per_launch": nf,
                     "one_frame_kernel_ms": kms, "composite_ms": cms, "clear_ms": zms,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "step_frac": (n * ALGO_BYTES_PER_POINT_STEP + w * h * ALGO_BYTES_PER_PIXEL_FRAME) / (ms / K * 1e-3) / 1e9 / hbm_peak},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    ctx.close()
    return 0


def _claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version
    line on stdout when NCCL_DEBUG=VERSION).  Point fd 1 at stderr for the duration of the run and keep
    the real stdout for the final line: everything `print`ed from here on goes through `sys.stdout`,
    which is re-opened on the saved descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
