"""advect_points across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

What shards and what does not (BASELINE.json north_star):

  points      advect independently        -> contiguous index ranges, one per rank; no communication
  velocity    read at random by every point -> replicated on every GPU
  splat       every rank counts ITS points into a private counter image (one cell per texel);
              the images are summed across ranks each frame and every rank composites (decay +
              replay + u8) the row band it owns.  The sum is a reduce-scatter by bands of padded
              rows over NCCL / NVLink, plus an all-gather of the 2*halo edge rows each band needs
              from its neighbours.
  frames      stay distributed: rank g holds image rows owned_rows(g); gather_image() assembles them.

The host-side plan (who owns which points and rows) is plain Python and is exercised on CPU with the
gloo backend in tests/test_multigpu_cpu.py; the device work goes through the C ABI
(clumpy_cuda_advect_count / _cells / _composite_rows / _finish_frame).
"""
import json
import os
import sys
import time

import numpy as np

CELL_ROW_MULTIPLE = 16  # composite tiles are 16 rows tall


def shard_range(n, world, rank):
    """Contiguous, balanced index range of rank `rank` among `world` for n items."""
    return n * rank // world, n * (rank + 1) // world


def shard_indices(pts, height, world, rank, how="home_rows"):
    """Indices (ascending) of the points rank `rank` simulates.  The shards partition range(len(pts)).

    "index":     contiguous index ranges -- no assumption about where points are.
    "home_rows": rank r takes the points whose HOME row lies in the r-th of `world` horizontal strips,
                 the strip edges being row quantiles of the cloud, so the shards are balanced (to within
                 the points of one image row) whatever the density.  Points stay within a few hundred
                 texels of home (they return there every nframes frames), so a rank's velocity gathers
                 and counter updates stay inside about 1/world of the image -- small enough to live in
                 L2 -- instead of paying a 32-byte sector of the whole field per point, and for a
                 roughly uniform cloud most of a rank's entries are for its own band."""
    n = len(pts)
    if how == "index" or world == 1 or n == 0:
        lo, hi = shard_range(n, world, rank)
        return np.arange(lo, hi)
    # same truncation as the splat (int32 cast); anything outside the image goes to the nearest row
    y = np.nan_to_num(pts[:, 1].astype(np.float64), nan=0.0, posinf=float(height), neginf=0.0)
    row = np.clip(y, 0, height - 1).astype(np.int64)
    below = np.cumsum(np.bincount(row, minlength=height))  # points with home row <= y
    # first row of strip r: the first row at which n*r/world points lie strictly above
    edges = [int(np.searchsorted(below, n * r // world, side="right")) if r else 0 for r in range(world)] + [height]
    return np.nonzero((row >= edges[rank]) & (row < edges[rank + 1]))[0]


def field_rows(height, world, rank):
    """Rows [r0, r1) of a (height, width) field that rank `rank` generates: equal bands (the last ranks
    one row shorter when `world` does not divide `height`)."""
    return shard_range(height, world, rank)


def sharded_simplex_curl(ctx, width, height, amplitude, frequency, seed, dist, device):
    """generate_simplex -> curl_2d with the ROWS sharded across the ranks (BASELINE configs[2]): rank r
    evaluates the potential on its band plus the one row below it -- curl_2d is a forward difference
    (curl_2d.cc:57-71), so the halo is one row in one direction, and recomputing it costs 1/band of the
    band instead of a point-to-point exchange -- curls the band, and the velocity bands are all-gathered
    so that every rank ends up with the whole field (advect_points replicates it).  Returns a
    (height, width, 2) float32 torch tensor on `device`, bit-identical to the single-GPU field."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    r0, r1 = field_rows(height, world, rank)
    halo = 1 if r1 < height else 0
    pot = torch.empty((max(r1 - r0 + halo, 1), width), dtype=torch.float32, device=device)
    band = torch.zeros((max(r1 - r0, 1), width, 2), dtype=torch.float32, device=device)
    if r1 > r0:
        ctx.generate_simplex_dev(width, height, r0, r1 - r0 + halo, amplitude, frequency, seed, pot.data_ptr())
        ctx.curl_2d_dev(pot.data_ptr(), width, r1 - r0, pot[r1 - r0].data_ptr() if halo else None, band.data_ptr())
    if world == 1:
        return band[:height]
    # bands differ by at most one row: gather them padded to the longest
    longest = max(field_rows(height, world, r)[1] - field_rows(height, world, r)[0] for r in range(world))
    padded = torch.zeros((longest, width, 2), dtype=torch.float32, device=device)
    padded[:r1 - r0].copy_(band[:r1 - r0])
    everything = torch.empty((world, longest, width, 2), dtype=torch.float32, device=device)
    if everything.is_cuda:
        dist.all_gather_into_tensor(everything, padded)
    else:
        dist.all_gather(list(everything.unbind(0)), padded)
    return torch.cat([everything[r, :field_rows(height, world, r)[1] - field_rows(height, world, r)[0]] for r in range(world)])


class BandPlan:
    """Row bands of the padded counter image (rows_alloc = world * chunk rows)."""

    def __init__(self, height, halo, rows_alloc, world):
        if rows_alloc % world:
            raise ValueError("the counter image must have a multiple of world_size rows")
        if rows_alloc < height + 2 * halo:
            raise ValueError("the counter image is too small for the framebuffer")
        self.height, self.halo, self.rows_alloc, self.world = height, halo, rows_alloc, world
        self.chunk = rows_alloc // world
        if world > 1 and self.chunk < 2 * halo:
            raise ValueError("bands thinner than the sprite are not supported")

    def padded_rows(self, rank):
        """Padded rows [lo, hi) that rank `rank` receives from the reduce-scatter."""
        return rank * self.chunk, (rank + 1) * self.chunk

    def owned_rows(self, rank):
        """Image rows [y0, y1) whose pixels rank `rank` composites: those whose whole window of
        2*halo+1 padded rows lies in the rank's chunk plus its two halos."""
        lo, hi = self.padded_rows(rank)
        y0, y1 = max(0, lo - self.halo), min(self.height, hi - self.halo)
        return (y0, max(y0, y1))

    def band_buffer_rows(self):
        """Rows of the per-rank band buffer: halo above + chunk + halo below + tile slack."""
        return self.chunk + 2 * self.halo + 2 * CELL_ROW_MULTIPLE

    def cell_row_offset(self, rank):
        """Row of the band buffer that holds padded row owned_rows(rank)[0] (= the cells of image row
        y0 - halo): what composite_rows wants a pointer to."""
        lo, _ = self.padded_rows(rank)
        y0, _ = self.owned_rows(rank)
        return y0 - (lo - self.halo)


def exchange_counts(private, band, plan, rank, dist, scratch=None):
    """Sums the private counter images of all ranks and leaves rank `rank`'s band (chunk + halos) in
    `band` (rows: halo | chunk | halo | slack).  `private` is (rows_alloc, pitch), `band` is
    (band_buffer_rows, pitch); both torch tensors on the rank's device (or CPU for the gloo tests)."""
    import torch
    h, chunk, world = plan.halo, plan.chunk, plan.world
    mid = band[h:h + chunk]
    if world == 1:
        mid.copy_(private[:chunk])
    elif private.is_cuda:
        dist.reduce_scatter_tensor(mid, private, op=dist.ReduceOp.SUM)
    else:  # gloo has no reduce-scatter: all-reduce a copy and keep our rows (tests only)
        tmp = private.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        mid.copy_(tmp[rank * chunk:(rank + 1) * chunk])
    if h == 0:
        return
    band[:h].zero_()
    band[h + chunk:h + chunk + h].zero_()
    if world == 1:
        return
    edges = torch.cat([mid[:h], mid[chunk - h:chunk]])  # my first and last h rows
    gathered = torch.empty((world,) + tuple(edges.shape), dtype=edges.dtype, device=edges.device) \
        if scratch is None else scratch
    dist.all_gather_into_tensor(gathered, edges) if private.is_cuda else \
        dist.all_gather(list(gathered.unbind(0)), edges)
    if rank > 0:
        band[:h].copy_(gathered[rank - 1, h:2 * h])          # previous band's last rows
    if rank < world - 1:
        band[h + chunk:h + chunk + h].copy_(gathered[rank + 1, :h])  # next band's first rows


class _DeviceView:
    """Lets torch alias library-owned device memory (no copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


_TYPESTR = {1: ("|u1", "uint8"), 2: ("<f2", "float16"), 4: ("<i4", "int32")}


class ShardedAdvect:
    """advect_points with the points of one run spread over torch.distributed ranks."""

    def __init__(self, ctx, pts, vel, step_size, kernel_size, decay, nframes, dist, device):
        import torch
        import clumpy_b200
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.ctx, self.device = ctx, device
        self.h_img, self.w_img = vel.shape[:2]
        n = len(pts)
        lo, hi = shard_range(n, self.world, self.rank)
        offsets = clumpy_b200.age_offsets(nframes, n)[lo:hi]  # offsets follow the GLOBAL point index
        # count-only cells (the exact-order cells carry local indices); rows padded so that every
        # rank gets a whole number of 16-row tiles
        os.environ["CLUMPY_CUDA_CELLS"] = "f16" if kernel_size <= 3 else "u32"
        os.environ["CLUMPY_CUDA_OVERLAP"] = "0"
        os.environ["CLUMPY_CUDA_CELL_ROWS_MULTIPLE"] = str(CELL_ROW_MULTIPLE * self.world)
        self.adv = ctx.advect(np.ascontiguousarray(pts[lo:hi]), vel, step_size, kernel_size, decay, nframes,
                              age_offset=offsets)
        self.npts_local = hi - lo
        ptr, nbytes, pitch, rows, halo, bpc = self.adv.cells()
        typestr, tname = _TYPESTR[bpc]
        self.plan = BandPlan(self.h_img, halo, rows, self.world)
        self.private = torch.as_tensor(_DeviceView(ptr, (rows, pitch), typestr), device=device)
        self.band = torch.zeros((self.plan.band_buffer_rows(), pitch), dtype=getattr(torch, tname), device=device)
        self.scratch = torch.empty((self.world, 2 * halo, pitch), dtype=self.band.dtype, device=device) if halo else None
        self.y0, self.y1 = self.plan.owned_rows(self.rank)
        self.band_ptr = self.band.data_ptr() + self.plan.cell_row_offset(self.rank) * pitch * bpc

    def step(self, count=1):
        for _ in range(count):
            self.adv.count()
            exchange_counts(self.private, self.band, self.plan, self.rank, self.dist, self.scratch)
            self.adv.composite_rows(self.band_ptr, self.y0, self.y1 - self.y0)
            self.adv.finish_frame()

    def band_image(self):
        """This rank's rows of the current framebuffer as a (rows, W) uint8 torch tensor (a view)."""
        img = self.torch.as_tensor(_DeviceView(self.adv.image_dev(), (self.h_img, self.w_img), "|u1"), device=self.device)
        return img[self.y0:self.y1]

    def gather_image(self):
        """The whole framebuffer on every rank (numpy), assembled from the bands."""
        torch, dist = self.torch, self.dist
        full = torch.zeros((self.h_img, self.w_img), dtype=torch.uint8, device=self.device)
        full[self.y0:self.y1].copy_(self.band_image())
        if self.world > 1:
            dist.all_reduce(full, op=dist.ReduceOp.SUM)  # bands are disjoint
        return full.cpu().numpy()

    def points(self):
        """All positions in the caller's order (numpy), gathered from the shards."""
        torch, dist = self.torch, self.dist
        mine = torch.from_numpy(self.adv.points()).to(self.device)
        if self.world == 1:
            return mine.cpu().numpy()
        sizes = [shard_range(self._n_total(), self.world, r) for r in range(self.world)]
        parts = [torch.empty((hi - lo, 2), dtype=torch.float32, device=self.device) for lo, hi in sizes]
        dist.all_gather(parts, mine)
        return torch.cat(parts).cpu().numpy()

    def _n_total(self):
        t = self.torch.tensor([self.npts_local], device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def close(self):
        self.adv.close()


class RoutedAdvect:
    """advect_points across ranks with the splat exchange FUSED into the advect kernel: the counter
    image is distributed by row bands and every rank's advect kernel stores each point's cell index
    straight into the owning rank's inbox through peer-mapped memory (CUDA IPC over NVLink); see
    include/clumpy_cuda.h "multi-GPU with the exchange fused into the advect kernel".  Per frame the
    only collective is an all-gather of world x world entry counts, which is also the frame fence."""

    def __init__(self, ctx, pts, vel, step_size, kernel_size, decay, nframes, dist, device):
        import torch
        import clumpy_b200
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # profiling aid: one process plays rank r of w alone (CLUMPY_ROUTE_LOOPBACK=1 CLUMPY_ROUTE_LOOPBACK_AS=r/w);
        # timing and traffic are a real rank's, the frames are not (see clumpy_cuda_advect_route_begin)
        self.loopback = self.world == 1 and os.environ.get("CLUMPY_ROUTE_LOOPBACK") == "1"
        if self.loopback:
            self.rank, self.world = (int(v) for v in os.environ.get("CLUMPY_ROUTE_LOOPBACK_AS", "3/8").split("/"))
        self.ctx, self.device = ctx, device
        self.h_img, self.w_img = vel.shape[:2]
        n = self.n_total = len(pts)
        how = os.environ.get("CLUMPY_MULTIGPU_SHARD", "home_rows")
        self.indices = shard_indices(pts, self.h_img, self.world, self.rank, how)
        offsets = clumpy_b200.age_offsets(nframes, n)[self.indices]  # offsets follow the GLOBAL point index
        os.environ["CLUMPY_CUDA_CELLS"] = "f16" if kernel_size <= 3 else "u32"  # count-only cells
        os.environ.pop("CLUMPY_CUDA_CELL_ROWS_MULTIPLE", None)  # (the dense variant's: bands here are the library's own)
        self.adv = ctx.advect(np.ascontiguousarray(pts[self.indices]), vel, step_size, kernel_size, decay, nframes,
                              age_offset=np.ascontiguousarray(offsets))
        sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
        sizes[self.rank] = len(self.indices)
        if self.world > 1 and not self.loopback:
            dist.all_reduce(sizes)
        self.shard_sizes = [int(v) for v in sizes.tolist()]
        capacity = max(1, max(self.shard_sizes))
        inbox, _ = self.adv.route_begin(self.rank, self.world, capacity)
        # exchange CUDA IPC handles of the inboxes and map the peers'
        if self.loopback:
            self.peer_ptrs = [inbox] * self.world
        else:
            handles = [None] * self.world
            dist.all_gather_object(handles, ctx.ipc_export(inbox))
            self.peer_ptrs = [inbox if r == self.rank else ctx.ipc_open(handles[r]) for r in range(self.world)]
        self.adv.route_peers(self.peer_ptrs)
        self.counts_all = torch.zeros((self.world, self.world), dtype=torch.int32, device=device)
        self.y0, self.y1 = self.adv.owned_rows()
        self._counts_view = {}
        dist.barrier()

    def step(self, count=1):
        torch, dist = self.torch, self.dist
        if os.environ.get("CLUMPY_MULTIGPU_FENCE", "device") == "device":
            # fence on the device (epoch flags in peer memory): one C call enqueues all the frames
            self.adv.route_steps(count)
            return
        for _ in range(count):  # fence = NCCL all-gather of the entry counts, one round trip per frame
            ptr = self.adv.route()
            mine = self._counts_view.get(ptr)
            if mine is None:
                mine = torch.as_tensor(_DeviceView(ptr, (self.world,), "<i4"), device=self.device)
                self._counts_view[ptr] = mine
            if self.world > 1:
                dist.all_gather_into_tensor(self.counts_all, mine)  # also the frame fence
            else:
                self.counts_all[0].copy_(mine)
            self.adv.drain(self.counts_all.data_ptr())

    band_image = ShardedAdvect.band_image
    gather_image = ShardedAdvect.gather_image

    def points(self):
        """All positions in the caller's order (numpy), gathered from the shards."""
        torch, dist = self.torch, self.dist
        mine = self.adv.points()
        if self.world == 1:
            return mine
        cap = max(self.shard_sizes)
        pos = torch.zeros((cap, 2), dtype=torch.float32, device=self.device)
        idx = torch.zeros((cap,), dtype=torch.int64, device=self.device)
        pos[:len(mine)] = torch.from_numpy(mine).to(self.device)
        idx[:len(mine)] = torch.from_numpy(self.indices).to(self.device)
        all_pos = torch.empty((self.world, cap, 2), dtype=torch.float32, device=self.device)
        all_idx = torch.empty((self.world, cap), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(all_pos, pos)
        dist.all_gather_into_tensor(all_idx, idx)
        out = np.empty((self.n_total, 2), dtype=np.float32)
        all_pos, all_idx = all_pos.cpu().numpy(), all_idx.cpu().numpy()
        for r, m in enumerate(self.shard_sizes):
            out[all_idx[r, :m]] = all_pos[r, :m]
        return out

    def close(self):
        self.torch.cuda.synchronize()
        timeouts = self.adv.route_errors()
        if self.world > 1 and not self.loopback:
            self.dist.barrier()  # nobody may still be writing into an inbox that is about to go away
        if timeouts:
            raise RuntimeError(f"rank {self.rank}: {timeouts} frame fences timed out (a peer stopped publishing)")
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank and p and not self.loopback:
                self.ctx.ipc_close(p)
        self.adv.close()


def make_sharded(kind, *a, **kw):
    """kind: "routed" (exchange fused into the advect kernel, default) or "dense" (counter images
    summed with an NCCL reduce-scatter)."""
    return (RoutedAdvect if kind == "routed" else ShardedAdvect)(*a, **kw)


def bench_multi(args, cfg, workload, metric, unit):
    """bench.py body for --gpus N > 1 (launched by torch.distributed.run, one rank per GPU)."""
    import torch
    import torch.distributed as dist
    import clumpy_b200
    from bench import ALGO_BYTES_PER_PIXEL_FRAME, ALGO_BYTES_PER_POINT_STEP, ClockSampler, measured_peaks, synth_points

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

    # points come from the same seeded recipe on every rank and are sliced by the sharded run
    # the (replicated) field: simplex + curl row-sharded across the ranks, bands all-gathered over NCCL
    d_vel = sharded_simplex_curl(ctx, w, h, cfg["amplitude"], cfg["frequency"] / rows_div, cfg["seed"], dist, device)
    vel = torch.empty((h, w, 2), dtype=torch.float32).pin_memory()
    vel.copy_(d_vel)
    torch.cuda.synchronize()
    del d_vel
    pts = synth_points(n, w, h)

    def note(msg):  # progress on stderr (CLUMPY_BENCH_VERBOSE=1)
        if os.environ.get("CLUMPY_BENCH_VERBOSE") == "1":
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:.2f}s] {msg}", file=sys.stderr, flush=True)

    kind = os.environ.get("CLUMPY_MULTIGPU", "routed")
    note("inputs ready")
    run = make_sharded(kind, ctx, pts, vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], cfg["nframes"],
                       dist, device)
    # CLUMPY_BENCH_NO_CLOCKS=1 (A/B runs): nvidia-smi polling GPU 0 every 50 ms is not free for rank 0
    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get("CLUMPY_BENCH_NO_CLOCKS") != "1" else None
    note("sharded run set up")
    run.step(W)
    note("warm-up enqueued")
    dist.barrier()
    torch.cuda.synchronize()
    note("warm-up done")
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run.step(K)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # slowest rank
    ms = float(ms.item())
    note(f"timed region done: {ms / K:.4f} ms per frame")
    launches = ctx.launch_count() - launches0
    value = n * K / (ms * 1e-3)

    # per-rank kernel times (CUDA events around every kernel; the drain includes the wait for the peers)
    kernel_ms = None
    if kind == "routed" and os.environ.get("CLUMPY_MULTIGPU_FENCE", "device") == "device":
        t = torch.tensor(run.adv.route_profile(50), device=device)
        allt = torch.empty((world, 3), device=device)
        dist.all_gather_into_tensor(allt, t)
        kernel_ms = {"route": allt[:, 0].tolist(), "drain_incl_wait": allt[:, 1].tolist(), "composite": allt[:, 2].tolist(),
                     "points_per_rank": run.shard_sizes}

    note(f"kernel profile done: {kernel_ms}")
    # e2e: every rank also copies its band of each recorded frame to pinned host memory
    y0, y1 = run.y0, run.y1
    host_band = [torch.empty((max(1, y1 - y0), w), dtype=torch.uint8).pin_memory() for _ in range(2)]
    nrec = max(1, K // 2)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run2 = make_sharded(kind, ctx, pts, vel.numpy(), cfg["step_size"], cfg["kernel_size"], cfg["decay"], nrec, dist, device)
    note("e2e: second run set up")
    run2.step(nrec)
    note("e2e: warm-up frames enqueued")
    for f in range(nrec):
        run2.step(1)
        if y1 > y0:
            host_band[f & 1][:y1 - y0].copy_(run2.band_image(), non_blocking=True)
        if f % 50 == 0:
            note(f"e2e: recorded frame {f} enqueued")
    note("e2e: all frames enqueued")
    torch.cuda.synchronize()
    note(f"e2e: device idle, fence time-outs so far: {run2.adv.route_errors() if kind == 'routed' else 0}")
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    run2.close()
    run.close()
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        e2e_steps = 2 * nrec
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": workload, "frames_per_s": K / (ms * 1e-3),
                       "parallelism": (f"points sharded x{world} by home row, field replicated, counter image distributed by "
                                       "row bands; a rank's route kernel (up to 4 frames per launch) counts the points "
                                       "that land in its own band directly and stores the others' cell indices into the "
                                       "band owners' inboxes through peer memory (NVLink); device-side frame fence (epoch "
                                       "flags in peer memory); drain + band composite per rank on side streams; no "
                                       "collective on the data path") if kind == "routed" else
                                      (f"points sharded x{world}, field replicated, per-frame NCCL reduce-scatter of the "
                                       "f16 counter image by row bands + halo all-gather, band composite per rank"),
                       "l2": "inputs larger than L2, no flush: per rank and ring period (8 frames) the frames touch positions + "
                             "phases (21 MB at 8 ranks), the rank's part of the field (~35-60 MB), 8 band buffers "
                             "(8 x 8.6 MB), the image band and the inbox slots -- ~160 MB at 8 ranks, more with fewer "
                             "ranks -- against 126 MB of L2"},
            "clocks": clocks,
            "e2e": {"value": n * e2e_steps / dt, "unit": unit,
                    "h2d_bytes_per_step": (n * 12 + world * w * h * 8) / e2e_steps, "d2h_bytes_per_step": nrec * w * h / e2e_steps,
                    "steps": e2e_steps, "seconds": dt,
                    "what": "per rank: H2D of its point shard + the whole field, sort, frames, D2H of its row band of "
                            "every recorded frame to pinned memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole step (all ranks)", "peak": hbm_peak * world, "unit": "GB/s",
                         "achieved": (n * ALGO_BYTES_PER_POINT_STEP + w * h * ALGO_BYTES_PER_PIXEL_FRAME) / (ms / K * 1e-3) / 1e9,
                         "frac": (n * ALGO_BYTES_PER_POINT_STEP + w * h * ALGO_BYTES_PER_PIXEL_FRAME) / (ms / K * 1e-3) / 1e9 / (hbm_peak * world),
                         "traffic": None, "peak_source": peak_src, "per_rank_kernel_ms": kernel_ms},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    ctx.close()
    dist.destroy_process_group()
    return 0
