#!/usr/bin/env python3
"""bench.py -- advect_points throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3]

Workload (BASELINE.json configs[3], "C4"): 16 777 216 uniform synthetic points
(numpy default_rng(0)) on an 8192 x 4096 curl-of-simplex velocity field
(generate_simplex 8192x4096 1.0 8.0 0 -> curl_2d), kernel_size 3, step_size 240, decay 0.99,
nframes 1000.  A "step" is ONE simulation frame of advect_points over all points: Euler step +
reset + splat + u8 decay/composite of the 8192 x 4096 framebuffer.

value     point-steps/s with everything resident in HBM, CUDA events on the launching stream.
steady_ms_per_step  the same over at least three regroup periods (the points are re-sorted by tile every 64
          frames; a 20-step window right after a sort is the best case).
e2e       the same metric through the public C-ABI call clumpy_cuda_advect_points with HOST (pinned)
          buffers: upload of points + field, reset offsets drawn on the host, K simulation frames, and a
          device->host copy of each of the K/2 recorded 32 MiB frames; wall clock, device idle on both sides.
roofline  both stages of the step (advect launch, composite launch) timed live with CUDA events around each
          launch, against the measured HBM copy bandwidth, and the splat's reductions against the measured
          atomic throughput (clumpy_cuda_atomic_peak, same run); `bound` is the slower leg of the whole step.
parity    the GPU frames against the reference CPU command's frames on the SAME inputs at full size (2 recorded
          frames, all 16.7 M points): fraction of pixels within 1 LSB and the largest difference.  Below 0.999
          the run fails.
cpu_baseline  the reference CLI (oracle/_ref/clumpy_ref, single-threaded like the reference) on the
          same inputs for a bounded number of frames, on this box's host cores.

--impl reference times only that CPU arm and prints the same line shape with "impl": "reference".
--workload c3 measures BASELINE.json configs[2] instead (4-octave generate_simplex + curl_2d at 16384 x 16384,
rows sharded across the GPUs); its line is an auxiliary measurement, not the headline metric.
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

# The advect pipeline keeps kernels on several streams in flight (route / drain + composite / frame copy)
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
PARITY_BAR = 0.999              # BASELINE.json: u8 frames within 1 LSB on >= 99.9 % of the pixels
# dram__bytes_read.sum + dram__bytes_write.sum per launch (8 frames, C4, steady state) from the round-2
# `ncu --set full` captures of the two kernels (profiles/r2_ncu_steady_summary.txt: 1.269 + 0.747 GB;
# profiles/r2_ncu_composite_lookahead_summary.txt: 0.573 + 0.528 GB).  Constants, not measured by this script: a
# number taken under a profiler is never a bench value.
NCU_TRAFFIC_BYTES = {"advect_fused_kernel": 2_016_050_000, "composite_group_kernel": 1_100_655_000}
C3 = dict(width=16384, height=16384, octaves=[(1.0, 8.0, 0), (0.5, 16.0, 1), (0.25, 32.0, 2), (0.125, 64.0, 3)])
C3_WORKLOAD = "C3: generate_simplex 16384x16384, 4 octaves (amp, freq, seed) = (1,8,0) (0.5,16,1) (0.25,32,2) (0.125,64,3) summed in FP32 in that order, then curl_2d; rows sharded across the GPUs"


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

def cpu_reference_run(pts, vel, cfg, nframes_cpu, keep_frames=False):
    """Runs `advect_points` of the compiled reference for nframes_cpu recorded frames (2*nframes_cpu simulation
    frames) over ALL points; returns (point-steps/s, kind, description, frames or None)."""
    import oracle
    n = len(pts)
    steps = 2 * nframes_cpu
    frames = None
    if oracle.have_ref():
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        tmp = tempfile.mkdtemp(prefix="clumpy_ref_", dir=base)
        try:
            np.save(os.path.join(tmp, "pts.npy"), pts)
            np.save(os.path.join(tmp, "vel.npy"), vel)
            t0 = time.perf_counter()
            oracle.ref_cli("advect_points", "pts.npy", "vel.npy", cfg["step_size"], cfg["kernel_size"],
                           cfg["decay"], nframes_cpu, "f.npy", cwd=tmp)
            dt = time.perf_counter() - t0
            if keep_frames:
                frames = np.stack([np.load(os.path.join(tmp, f"{f:03d}f.npy")) for f in range(nframes_cpu)])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        kind = "reference"
        what = (f"oracle/_ref/clumpy_ref advect_points (unmodified reference, -O3, 1 thread), whole command incl. "
                f".npy load/save on tmpfs, all {n} points, nframes={nframes_cpu} ({steps} simulation frames)")
    else:
        oracle.lib()
        a = oracle.Advect(pts, vel, cfg["step_size"], cfg["kernel_size"], cfg["decay"], nframes_cpu)
        got = []
        t0 = time.perf_counter()
        for s in range(steps):
            a.step()
            if keep_frames and s >= nframes_cpu:
                got.append(a.image())
        dt = time.perf_counter() - t0
        a.close()
        frames = np.stack(got) if keep_frames else None
        kind = "port"
        what = (f"oracle/liboracle.so restatement (the reference could not be compiled here), 1 thread, compute only, "
                f"all {n} points, {steps} simulation frames")
    return n * steps / dt, kind, what, frames


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
    if args.workload == "c3":
        return run_reference_c3(args)
    cfg = dict(C4)
    nframes_cpu = max(1, min(3, args.steps // 2))
    pts, vel = host_inputs_for_reference(cfg)
    value, kind, what, _ = cpu_reference_run(pts, vel, cfg, nframes_cpu)
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
# GPU arm, one GPU
# ---------------------------------------------------------------------------------------------

def roofline_block(n, w, h, ms_per_step, stages_live, stages_alone, nf, hbm_peak, peak_src, atomic_gups, world=1):
    """The roofline of the whole step and of its two stages.  Legs of a step: HBM (28 B per point-step + 2 B per
    pixel and frame) and the splat's reductions (one per point-step) against the measured atomic
    throughput; the step's bound is the slower leg, its fraction that leg's time over the measured step time."""
    algo_step = n * ALGO_BYTES_PER_POINT_STEP + w * h * ALGO_BYTES_PER_PIXEL_FRAME
    hbm_leg_ms = algo_step / (hbm_peak * world * 1e9) * 1e3
    atomic_leg_ms = n / (atomic_gups * world * 1e9) * 1e3 if atomic_gups else None
    bound = "hbm" if atomic_leg_ms is None or hbm_leg_ms >= atomic_leg_ms else "l2-atomic"
    leg_ms = hbm_leg_ms if bound == "hbm" else atomic_leg_ms
    out = {
        "bound": bound, "kernel": "whole step: advect_fused_kernel + composite_group_kernel" if world == 1 else "whole step (all ranks)",
        "achieved": algo_step / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
        "frac": leg_ms / ms_per_step, "hbm_frac": hbm_leg_ms / ms_per_step,
        "atomic_frac": atomic_leg_ms / ms_per_step if atomic_leg_ms else None,
        "atomic_peak_G_updates_per_s": atomic_gups,
        "atomic_peak_source": "clumpy_cuda_atomic_peak in this run: red.global.add.noftz.f16x2, tile-ordered updates on one 67 MB counter image",
        "peak_source": peak_src, "algorithmic_bytes_per_step": algo_step, "traffic": None,
    }
    if stages_live:
        adv_ms, comp_ms = stages_live
        adv_alone, comp_alone = stages_alone if stages_alone else (None, None)
        adv_bytes, comp_bytes = n * ALGO_BYTES_PER_POINT_STEP * nf, w * h * ALGO_BYTES_PER_PIXEL_FRAME * nf
        out["frames_per_launch"] = nf
        out["stages"] = {
            "advect_fused_kernel": {
                "launch_ms": adv_ms, "launch_ms_alone": adv_alone, "algorithmic_bytes_per_launch": adv_bytes,
                "achieved": adv_bytes / (adv_ms * 1e-3) / 1e9, "frac": adv_bytes / (adv_ms * 1e-3) / 1e9 / hbm_peak,
                "frac_alone": adv_bytes / (adv_alone * 1e-3) / 1e9 / hbm_peak if adv_alone else None,
                "atomic_frac": (n * nf / (adv_ms * 1e-3) / 1e9) / atomic_gups if atomic_gups else None,
                "traffic": NCU_TRAFFIC_BYTES["advect_fused_kernel"]},
            "composite_group_kernel": {
                "launch_ms": comp_ms, "launch_ms_alone": comp_alone, "algorithmic_bytes_per_launch": comp_bytes,
                "achieved": comp_bytes / (comp_ms * 1e-3) / 1e9, "frac": comp_bytes / (comp_ms * 1e-3) / 1e9 / hbm_peak,
                "frac_alone": comp_bytes / (comp_alone * 1e-3) / 1e9 / hbm_peak if comp_alone else None,
                "traffic": NCU_TRAFFIC_BYTES["composite_group_kernel"]},
        }
        dominant = "advect_fused_kernel" if adv_ms >= comp_ms else "composite_group_kernel"
        out["dominant_stage"] = dominant
        out["traffic"] = out["stages"][dominant]["traffic"]
        out["traffic_source"] = "ncu --set full capture of the dominant stage's kernel on this workload, per launch (profiles/)"
    return out


def run_ours(args):
    import torch
    import clumpy_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if args.workload == "c3":
        return run_c3(args)
    if world > 1:
        return run_multi(args)

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

    # ---- steady state: at least three regroup periods, wherever the run is in its regroup cycle ----
    S = max(K, 200) if K < 2000 else K
    if S != K:
        e0.record(stream)
        adv.step(S)
        e1.record(stream)
        torch.cuda.synchronize()
        steady_ms = e0.elapsed_time(e1) / S
    else:
        steady_ms = ms / K

    # ---- roofline: both stages, live in the pipeline and alone; the atomic leg -------------------
    adv_ms, comp_ms, nf = adv.profile_stages(24)
    adv_alone, comp_alone, _ = adv.profile_stages(24, serial=True)
    atomic_ms = ctx.atomic_peak(0, 2, (w + 8) * (h + 2) * 2 // 256 * 256, n, iters=10)
    atomic_gups = n / (atomic_ms * 1e-3) / 1e9
    adv.close()
    roof = roofline_block(n, w, h, ms / K, (adv_ms, comp_ms), (adv_alone, comp_alone), nf, hbm_peak, peak_src, atomic_gups)
    roof["steady_frac"] = roof["frac"] * (ms / K) / steady_ms

    # The clock sampler stops here: it covers the device-timed regions above.  The e2e leg below is host wall clock,
    # and an nvidia-smi query every 50 ms takes driver locks that its allocations and launches also want (one run in
    # six measured 183 ms instead of 15.6 with the sampler still polling).
    clocks = sampler.stop()

    # ---- e2e: whole command through the C ABI with host buffers ---------------------------------
    nrec = max(1, K // 2)
    ctx.advect_points(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], 2,
                      on_frame=lambda f, px: 0)  # untimed warm-up call: module load, memory pool, pinned staging
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.advect_points(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], nrec,
                      on_frame=lambda f, px: 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e_steps = 2 * nrec
    e2e_value = n * e2e_steps / dt
    h2d = (n * 8 + w * h * 8) / e2e_steps  # points + field (the reset offsets are drawn on the device)
    d2h = (nrec * w * h) / e2e_steps

    # ---- CPU baseline + parity at full size: the reference on the same inputs, 2 recorded frames ----
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        v, kind, what, ref_frames = cpu_reference_run(pts.numpy(), vel.numpy(), cfg, nframes_cpu=2, keep_frames=True)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": what,
               "host_cpus": os.cpu_count()}
        gpu_frames = ctx.advect_points(pts.numpy(), vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], 2)
        d = np.abs(gpu_frames.astype(np.int16) - ref_frames.astype(np.int16))
        parity = {"within1": float(np.mean(d <= 1)), "max": int(d.max()), "identical": float(np.mean(d == 0)),
                  "frames": int(len(ref_frames)), "pixels_per_frame": int(w * h), "bar": PARITY_BAR,
                  "against": kind, "what": "the 2 recorded frames of advect_points with nframes = 2 on the full C4 inputs "
                                           "(all 16 777 216 points, 8192 x 4096), GPU vs the CPU arm's own output files"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "steady_ms_per_step": steady_ms, "steady_steps": S,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32+u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2, no flush (per group of 8 frames: positions 134 MB + "
                   "phases 34 MB + field 268 MB + 8 counter images of 67 MB written, read and zeroed vs 126 MB L2)",
                   "frames_per_s": K / (ms * 1e-3), "parallelism": "1 GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "seconds": dt,
                "what": "clumpy_cuda_advect_points with pinned host buffers: H2D of points + field, reset offsets drawn "
                        f"on the device, sort, {e2e_steps} simulation frames, D2H of {nrec} recorded frames; "
                        "after one untimed call of the same command (2 recorded frames)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "parity": parity,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    ctx.close()
    if parity and parity["within1"] < PARITY_BAR:
        print(f"PARITY FAILURE: only {parity['within1']:.6f} of the pixels within 1 LSB of the reference", file=sys.stderr)
        return 1
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm, several GPUs (launched by torch.distributed.run, one rank per GPU)
# ---------------------------------------------------------------------------------------------

def run_multi(args):
    import torch
    import torch.distributed as dist
    import clumpy_b200
    import clumpy_b200.multigpu as mg

    cfg, workload = dict(C4), WORKLOAD
    t_start = time.perf_counter()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = clumpy_b200.Context(local_rank, stream.cuda_stream)
    n, w, h = cfg["npts"], cfg["width"], cfg["height"]
    # development aid (never set by the driver): CLUMPY_BENCH_ROWS_DIV=d runs 1/d of the workload --
    # the top h/d rows of the field and n/d points -- so that `world` ranks see the per-rank working
    # set of a (world * d)-rank run of the full workload.  The JSON line says so in config.workload.
    rows_div = int(os.environ.get("CLUMPY_BENCH_ROWS_DIV", "1"))
    if rows_div > 1:
        n, h = n // rows_div, h // rows_div
        workload = f"SCALED 1/{rows_div} (development run, not the benchmark): {n} points on {w}x{h}; " + workload
    K, W = args.steps, args.warmup
    hbm_peak, peak_src = measured_peaks()

    def note(msg):  # progress on stderr (CLUMPY_BENCH_VERBOSE=1)
        if os.environ.get("CLUMPY_BENCH_VERBOSE") == "1":
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:.2f}s] {msg}", file=sys.stderr, flush=True)

    # the (replicated) field: simplex + curl row-sharded across the ranks, bands all-gathered over NCCL; it
    # stays on the device and the runs use it in place.  A pinned host copy of this rank's band feeds the e2e leg.
    d_vel = mg.sharded_simplex_curl(ctx, w, h, cfg["amplitude"], cfg["frequency"] / rows_div, cfg["seed"], dist, device).contiguous()
    r0, r1 = mg.field_rows(h, world, rank)
    vel_band = torch.empty((max(r1 - r0, 1), w, 2), dtype=torch.float32).pin_memory()
    vel_band[:r1 - r0].copy_(d_vel[r0:r1])
    # points come from the same seeded recipe on every rank; the rank's shard is selected on the device
    pts = torch.from_numpy(synth_points(n, w, h)).pin_memory()
    torch.cuda.synchronize()
    note("inputs ready")

    def make_run(points, field, nframes):
        return mg.RoutedAdvect(ctx, points, field, cfg["step_size"], cfg["kernel_size"], cfg["decay"], nframes, dist, device)

    run = make_run(pts.numpy(), d_vel, cfg["nframes"])
    # CLUMPY_BENCH_NO_CLOCKS=1 (A/B runs): nvidia-smi polling GPU 0 every 50 ms is not free for rank 0
    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get("CLUMPY_BENCH_NO_CLOCKS") != "1" else None
    note("sharded run set up")
    run.step(W)
    dist.barrier()
    torch.cuda.synchronize()
    note("warm-up done")

    def timed(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        run.step(steps)
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # slowest rank
        return float(t.item())

    launches0 = ctx.launch_count()
    ms = timed(K)
    launches = ctx.launch_count() - launches0
    note(f"timed region done: {ms / K:.4f} ms per frame")
    value = n * K / (ms * 1e-3)
    S = max(K, 200) if K < 2000 else K
    steady_ms = timed(S) / S if S != K else ms / K

    # per-rank kernel times (CUDA events around every kernel; the drain includes the wait for the peers)
    rt, dr, co, nf = run.adv.route_profile(12)
    t = torch.tensor([rt / nf, dr / nf, co / nf], device=device)
    allt = torch.empty((world, 3), device=device)
    dist.all_gather_into_tensor(allt, t)
    kernel_ms = {"frames_per_group": nf, "route_per_frame": allt[:, 0].tolist(), "drain_incl_wait_per_frame": allt[:, 1].tolist(),
                 "composite_per_frame": allt[:, 2].tolist(), "points_per_rank": run.shard_sizes}
    note(f"kernel profile done: {kernel_ms}")
    atomic_ms = ctx.atomic_peak(0, 2, (w + 8) * (h + 2) * 2 // 256 * 256, n, iters=5)
    atomic_gups = n / (atomic_ms * 1e-3) / 1e9
    run.close()

    clocks = sampler.stop() if sampler else None  # (the device-timed regions are over; see run_ours)

    # ---- e2e: host buffers in, this rank's band of every recorded frame out ------------------------
    nrec = max(1, K // 2)
    y0, y1 = mg.field_rows(h, world, rank)  # (an upper bound of the band height for the pinned buffer)
    band_rows = -(-h // world) + 32

    def e2e_once(nr, host_frames):
        t_e2e = time.perf_counter()

        def lap(what):  # CLUMPY_BENCH_VERBOSE=1: where the end-to-end time goes (synchronises: perturbs the run)
            if os.environ.get("CLUMPY_BENCH_VERBOSE") == "1":
                torch.cuda.synchronize()
                note(f"e2e +{(time.perf_counter() - t_e2e) * 1e3:8.2f} ms  {what}")

        # H2D: this rank's 1/world of the point list and its rows of the field; both are then all-gathered over NVLink
        # (every rank needs the whole list to select its strip, and the whole field)
        if n % world == 0:
            d_slice = pts[rank * (n // world):(rank + 1) * (n // world)].to(device, non_blocking=True)
            d_pts = torch.empty((n, 2), dtype=torch.float32, device=device)
            dist.all_gather_into_tensor(d_pts, d_slice)
        else:
            d_pts = pts.to(device, non_blocking=True)
        d_band = vel_band.to(device, non_blocking=True)
        field = torch.empty((world, vel_band.shape[0], w, 2), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(field, d_band)                  # NVLink: every rank gets the whole field
        field = field.reshape(-1, w, 2) if h % world == 0 else torch.cat(
            [field[r, :mg.field_rows(h, world, r)[1] - mg.field_rows(h, world, r)[0]] for r in range(world)])
        lap("points + field band uploaded, field all-gathered")
        r2 = make_run(d_pts, field, nr)                              # reset offsets, shard select, sort, IPC handles
        lap("sharded run set up")
        r2.step(nr)                                                  # warm-up half
        r2.record(nr, host_frames, host_frames.shape[1] * w)         # recorded half: D2H of the band of every frame
        torch.cuda.synchronize()
        r2.adv.wait_frames()
        lap("frames done")
        r2.close()
        lap("closed")

    host_frames = torch.empty((nrec, band_rows, w), dtype=torch.uint8).pin_memory()
    e2e_once(min(2, nrec), host_frames)  # untimed warm-up call (module load, memory pool, pinned staging, NCCL buffers)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_once(nrec, host_frames)
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    note(f"e2e done: {dt:.3f} s")
    if rank == 0:
        e2e_steps = 2 * nrec
        roof = roofline_block(n, w, h, ms / K, None, None, nf, hbm_peak, peak_src, atomic_gups, world=world)
        roof["per_rank_kernel_ms"] = kernel_ms
        roof["steady_frac"] = roof["frac"] * (ms / K) / steady_ms
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "steady_ms_per_step": steady_ms, "steady_steps": S,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": workload, "frames_per_s": K / (ms * 1e-3),
                       "parallelism": (f"points sharded x{world} by home row (strips of equal point count, selected on the device), "
                                       "field replicated (generated row-sharded, all-gathered over NVLink), counter image distributed by "
                                       "row bands; per group of 8 frames a rank's route kernel counts the points that land in its own band "
                                       "directly and stores the others' cell indices into the band owners' inboxes through peer memory "
                                       "(NVLink); one device-side fence per group (epoch flags in peer memory); drain + group composite of the "
                                       "band on a side stream; no collective on the data path"),
                       "l2": "inputs larger than L2, no flush: per rank and group of 8 frames the kernels touch the rank's positions + "
                             "phases, its part of the field, 16 band buffers and the image band -- more than the 126 MB of L2 at every N"},
            "clocks": clocks,
            "e2e": {"value": n * e2e_steps / dt, "unit": UNIT,
                    "h2d_bytes_per_step": ((1 if n % world == 0 else world) * n * 8 + w * h * 8) / e2e_steps, "d2h_bytes_per_step": nrec * w * h / e2e_steps,
                    "steps": e2e_steps, "seconds": dt,
                    "what": "per rank, from pinned host buffers: H2D of its 1/N of the point list + its rows of the field, both all-gathered "
                            "over NVLink, row histogram + shard selection on the device, reset offsets drawn on the device, sort, inbox "
                            "set-up + IPC handle exchange, frames, D2H of its row band of every recorded frame; after one untimed call of the same"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    ctx.close()
    dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# C3: multi-octave generate_simplex + curl_2d, rows sharded across the GPUs (auxiliary workload)
# ---------------------------------------------------------------------------------------------

def run_c3(args):
    import torch
    import clumpy_b200
    import clumpy_b200.multigpu as mg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = clumpy_b200.Context(local_rank, stream.cuda_stream)
    w, h, octaves = C3["width"], C3["height"], C3["octaves"]
    hbm_peak, peak_src = measured_peaks()
    K, W = min(args.steps, 50), max(3, min(args.warmup, 5))
    r0, r1 = mg.field_rows(h, world, rank)
    halo = 1 if r1 < h else 0  # curl_2d is a forward difference: one more row of the potential, recomputed here
    pot = torch.empty((r1 - r0 + halo, w), dtype=torch.float32, device=device)
    vel = torch.empty((r1 - r0, w, 2), dtype=torch.float32, device=device)
    amps, freqs, seeds = [o[0] for o in octaves], [o[1] for o in octaves], [o[2] for o in octaves]

    def step():
        ctx.generate_simplex_octaves_dev(w, h, r0, r1 - r0 + halo, amps, freqs, seeds, pot.data_ptr())
        ctx.curl_2d_dev(pot.data_ptr(), w, r1 - r0, pot[r1 - r0].data_ptr() if halo else None, vel.data_ptr())

    def sync_all():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(W):
        step()
    sync_all()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=device)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    launches = ctx.launch_count() - launches0
    # the two kernels one by one on this rank's band
    ctx.timer_start()
    for _ in range(K):
        ctx.generate_simplex_octaves_dev(w, h, r0, r1 - r0 + halo, amps, freqs, seeds, pot.data_ptr())
    simplex_ms = ctx.timer_stop() / K
    ctx.timer_start()
    for _ in range(K):
        ctx.curl_2d_dev(pot.data_ptr(), w, r1 - r0, pot[r1 - r0].data_ptr() if halo else None, vel.data_ptr())
    curl_ms = ctx.timer_stop() / K
    # e2e: the band of the velocity field back on the host (what `clumpy curl_2d` writes to disk), every step
    host = torch.empty((r1 - r0, w, 2), dtype=torch.float32).pin_memory()
    ke = max(2, K // 10)
    step(); host.copy_(vel, non_blocking=True); sync_all()
    t0 = time.perf_counter()
    for _ in range(ke):
        step()
        host.copy_(vel, non_blocking=True)
    sync_all()
    dt = time.perf_counter() - t0
    dtt = torch.tensor([dt], device=device)
    if dist:
        dist.all_reduce(dtt, op=dist.ReduceOp.MAX)
    dt = float(dtt.item())
    clocks = sampler.stop() if sampler else None
    # parity on a slice: rows of this rank's band against the CPU restatement of the multi-octave sum
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        import oracle
        nrows = 8
        want = None
        for a, f, s in octaves:
            o, _, _ = oracle.generate_simplex_rows(w, h, 0, nrows, a, f, s)
            want = o if want is None else (want + o).astype(np.float32)
        got = pot[:nrows].cpu().numpy()
        parity = {"rows": nrows, "max_abs_err": float(np.abs(got - want).max()), "bar": 1e-5 * sum(amps)}
    if rank == 0:
        px = w * h
        noct = len(octaves)
        line = {
            "metric": "generate_simplex (4 octaves) + curl_2d Mpx/s", "value": px / (ms * 1e-3) / 1e6, "unit": "Mpx/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 lattice placement + f32 contributions", "data": "synthetic",
            "config": {"workload": C3_WORKLOAD, "l2": "outputs larger than L2 (1.07 GB potential + 2.15 GB velocity over all ranks)",
                       "parallelism": f"rows sharded x{world}; the one halo row curl_2d needs is recomputed by the band above instead of exchanged"},
            "clocks": clocks,
            "e2e": {"value": px / (dt / ke) / 1e6, "unit": "Mpx/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": px * 8 / world, "steps": ke, "seconds": dt,
                    "what": "per step: the kernels + D2H of this rank's band of the velocity field into pinned memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32-issue (simplex) / hbm (curl)", "kernel": "simplex_octaves_kernel + curl_pair_kernel",
                         "simplex_ms": simplex_ms, "curl_ms": curl_ms,
                         "curl_achieved": px / world * 12 / (curl_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": px / world * 12 / (curl_ms * 1e-3) / 1e9 / hbm_peak,
                         "simplex_nominal_TFLOPs": px / world * 85 * noct / (simplex_ms * 1e-3) / 1e12,
                         "simplex_written_GBps": px / world * 4 / (simplex_ms * 1e-3) / 1e9,
                         "peak_source": peak_src, "traffic": None},
            "parity": parity,
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    ctx.close()
    if dist:
        dist.destroy_process_group()
    return 0


def run_reference_c3(args):
    """The reference CLI on the C3 inputs: 4 x generate_simplex + numpy sum + curl_2d on a bounded band."""
    import oracle
    w, h = C3["width"], 512  # a 512-row sample of the 16384 rows: the per-pixel cost does not depend on the row
    t0 = time.perf_counter()
    tmp = tempfile.mkdtemp(prefix="clumpy_c3_")
    try:
        total = None
        for i, (a, f, s) in enumerate(C3["octaves"]):
            if oracle.have_ref():
                oracle.ref_cli("generate_simplex", f"{w}x{h}", a, f * h / C3["height"], s, f"o{i}.npy", cwd=tmp)
                o = np.load(os.path.join(tmp, f"o{i}.npy"))
            else:
                o, _, _ = oracle.generate_simplex(w, h, a, f * h / C3["height"], s)
            total = o if total is None else total + o
        if oracle.have_ref():
            np.save(os.path.join(tmp, "pot.npy"), total)
            oracle.ref_cli("curl_2d", "pot.npy", "vel.npy", cwd=tmp)
        else:
            oracle.curl_2d(total)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    dt = time.perf_counter() - t0
    v = w * h / dt / 1e6
    kind = "reference" if oracle.have_ref() else "port"
    line = {"impl": "reference", "metric": "generate_simplex (4 octaves) + curl_2d Mpx/s", "value": v, "unit": "Mpx/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": C3_WORKLOAD, "cpu_rows": h},
            "cpu_baseline": {"value": v, "unit": "Mpx/s", "cores": 1, "kind": kind,
                             "sample": f"{w}x{h} sample: 4 generate_simplex commands + numpy sum + curl_2d, incl. .npy I/O"},
            "e2e": {"value": v, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
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
    ap.add_argument("--workload", default="c4", choices=["c4", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
