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
    edges = strip_edges(np.bincount(row, minlength=height), world)
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
        y0, y1 = min(self.height, max(0, lo - self.halo)), min(self.height, hi - self.halo)
        return (y0, max(y0, y1))  # (empty for a rank whose chunk lies wholly in the row padding)

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


def strip_edges(rows, world):
    """Row strips of equal point count: edges[r] is the first row of strip r (edges[world] = height), from the
    per-row point counts `rows` (Context.row_histogram or numpy.bincount)."""
    height, n = len(rows), int(np.sum(rows, dtype=np.int64))
    below = np.cumsum(rows, dtype=np.int64)  # points with home row <= y
    return [int(np.searchsorted(below, n * r // world, side="right")) if r else 0 for r in range(world)] + [height]


class ShardedAdvect:
    """advect_points with the points of one run spread over torch.distributed ranks, the counter images summed
    with an NCCL reduce-scatter every frame (the north star's literal design; kept as the baseline the routed
    variant is measured against)."""

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
        # count-only cells (the exact-order cells carry local indices); one stream; rows padded so that every
        # rank gets a whole number of 16-row tiles
        self.adv = ctx.advect(np.ascontiguousarray(pts[lo:hi]), vel, step_size, kernel_size, decay, nframes,
                              age_offset=offsets, cells="count", overlap=-1,
                              cell_rows_multiple=CELL_ROW_MULTIPLE * self.world)
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
    """advect_points across ranks with the splat exchange FUSED into the advect kernel: the counter image is
    distributed by row bands and every rank's route kernel stores each point's cell index straight into the
    owning rank's inbox through peer-mapped memory (CUDA IPC over NVLink); see include/clumpy_cuda.h "multi-GPU
    with the exchange fused into the advect kernel".  No collective on the data path.

    pts: (N, 2) numpy array (ALL the points; the rank's shard is selected on the device) or a torch CUDA tensor.
    vel: (H, W, 2) numpy array, or a torch CUDA tensor holding the whole (replicated) field, used in place."""

    def __init__(self, ctx, pts, vel, step_size, kernel_size, decay, nframes, dist, device, shard="home_rows"):
        import torch
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # profiling aid: one process plays rank r of w alone (CLUMPY_ROUTE_LOOPBACK=1 CLUMPY_ROUTE_LOOPBACK_AS=r/w);
        # timing and traffic are a real rank's, the frames are not (see clumpy_cuda_advect_route_begin)
        self.loopback = self.world == 1 and os.environ.get("CLUMPY_ROUTE_LOOPBACK") == "1"
        if self.loopback:
            self.rank, self.world = (int(v) for v in os.environ.get("CLUMPY_ROUTE_LOOPBACK_AS", "3/8").split("/"))
        self.ctx, self.device = ctx, device
        self.h_img, self.w_img = int(vel.shape[0]), int(vel.shape[1])
        self.n_total = n = int(pts.shape[0])
        d_pts = pts if torch.is_tensor(pts) else torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(device, non_blocking=True)
        self._vel = vel  # a device field is used in place: keep it alive
        torch.cuda.current_stream().synchronize()
        self._t0 = time.perf_counter()
        if shard == "home_rows":
            # strips of equal point count, from the row histogram of the whole cloud: every rank computes the
            # same edges from the same points, so nothing is exchanged
            rows = ctx.row_histogram(d_pts.data_ptr(), self.h_img, npts=n)
            edges = strip_edges(rows, self.world)
            self.shard_sizes = [int(rows[edges[r]:edges[r + 1]].sum()) for r in range(self.world)]
            lo, hi = edges[self.rank], edges[self.rank + 1]
            points_arg, npts_arg, rows_arg = d_pts.data_ptr(), n, ((lo, hi) if hi > lo else (self.h_img, self.h_img + 1))
            offsets = None  # computed by the library for the whole list while it uploads
        else:  # contiguous index ranges (tests: everything travels through the inboxes)
            import clumpy_b200
            lo, hi = shard_range(n, self.world, self.rank)
            self.shard_sizes = [shard_range(n, self.world, r)[1] - shard_range(n, self.world, r)[0] for r in range(self.world)]
            points_arg, npts_arg, rows_arg = d_pts[lo:hi].contiguous().data_ptr(), hi - lo, None
            self._index_base = lo
            offsets = np.ascontiguousarray(clumpy_b200.age_offsets(nframes, n)[lo:hi])
        vel_arg = vel.data_ptr() if torch.is_tensor(vel) else vel
        self.adv = ctx.advect(points_arg, vel_arg, step_size, kernel_size, decay, nframes, age_offset=offsets,
                              npts=npts_arg, width=self.w_img, height=self.h_img, cells="count", shard_rows=rows_arg)
        self.sharded_by_rows = rows_arg is not None
        self._lap("advect_begin_ex (uploads, offsets, shard selection, sort)")
        capacity = max(1, max(self.shard_sizes))
        # The inbox arena and the peers' mappings of it are kept in the context across runs (like a communicator):
        # exporting, exchanging and mapping a 2 GB allocation costs ~20 ms, more than a short run.  Every rank takes
        # the same decision: the size depends on world and capacity only.
        import clumpy_b200
        need = clumpy_b200.route_inbox_bytes(capacity, self.world)
        if self.loopback or self.world == 1:
            inbox, _ = self.adv.route_begin(self.rank, self.world, capacity)
            self.peer_ptrs = [inbox] * self.world
        else:
            cache = getattr(ctx, "_route_arena", None)
            if not (cache and cache["world"] == self.world and cache["bytes"] >= need):
                _drop_route_arena(ctx, dist)
                arena = ctx.dev_alloc(need)
                # exchange CUDA IPC handles of the arenas (64 bytes each, over NCCL) and map the peers'
                mine = torch.frombuffer(bytearray(ctx.ipc_export(arena)), dtype=torch.uint8).to(device)
                handles = torch.empty((self.world, 64), dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(handles, mine)
                handles = handles.cpu().numpy()
                peers = [arena if r == self.rank else ctx.ipc_open(handles[r].tobytes()) for r in range(self.world)]
                cache = ctx._route_arena = {"world": self.world, "inbox": arena, "bytes": need, "peers": peers}
            self.adv.route_begin(self.rank, self.world, capacity, arena=(cache["inbox"], cache["bytes"]))
            self.peer_ptrs = cache["peers"]
        self.adv.route_peers(self.peer_ptrs)
        self.y0, self.y1 = self.adv.owned_rows()
        if self.world > 1 and not self.loopback:
            dist.barrier()
        self._lap("inbox allocated, IPC handles exchanged and mapped, barrier")

    def _lap(self, what):  # CLUMPY_BENCH_VERBOSE=1: set-up phases on stderr
        if os.environ.get("CLUMPY_BENCH_VERBOSE") == "1":
            self.torch.cuda.synchronize()
            print(f"[RoutedAdvect rank {self.rank}] +{(time.perf_counter() - self._t0) * 1e3:8.2f} ms  {what}", file=sys.stderr, flush=True)

    def step(self, count=1):
        self.adv.route_steps(count)  # fence on the device (epoch flags in peer memory): one call enqueues all the frames

    def record(self, count, dst, frame_stride):
        """step(count); this rank's rows of every frame are copied to the pinned tensor / array `dst`."""
        self.adv.route_record(count, dst.data_ptr() if self.torch.is_tensor(dst) else dst, frame_stride)

    band_image = ShardedAdvect.band_image
    gather_image = ShardedAdvect.gather_image

    def points(self):
        """All positions in the caller's order (numpy), gathered from the shards."""
        torch, dist = self.torch, self.dist
        if self.sharded_by_rows:
            mine, index = self.adv.points_indexed()
        else:
            mine = self.adv.points()
            index = np.arange(self._index_base, self._index_base + len(mine), dtype=np.uint32)
        if self.world == 1 or self.loopback:
            out = np.empty((self.n_total, 2), dtype=np.float32)
            out[index] = mine
            return out
        cap = max(self.shard_sizes)
        pos = torch.zeros((cap, 2), dtype=torch.float32, device=self.device)
        idx = torch.zeros((cap,), dtype=torch.int64, device=self.device)
        pos[:len(mine)] = torch.from_numpy(mine).to(self.device)
        idx[:len(mine)] = torch.from_numpy(index.astype(np.int64)).to(self.device)
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
            self.dist.barrier()  # nobody may still be writing into an inbox that the next run re-uses
        self.adv.close()
        if timeouts:
            raise RuntimeError(f"rank {self.rank}: {timeouts} frame fences timed out (a peer stopped publishing)")


def _drop_route_arena(ctx, dist):
    """Unmaps and frees the inbox arena a context keeps between routed runs (collective: every rank decides alike)."""
    cache = getattr(ctx, "_route_arena", None)
    if not cache:
        return
    dist.barrier()
    for p in cache["peers"]:
        if p != cache["inbox"]:
            ctx.ipc_close(p)
    dist.barrier()  # every mapping of my arena is gone before I free it
    ctx.dev_free(cache["inbox"])
    ctx._route_arena = None


def make_sharded(kind, *a, **kw):
    """kind: "routed" (exchange fused into the advect kernel, default) or "dense" (counter images
    summed with an NCCL reduce-scatter)."""
    return (RoutedAdvect if kind == "routed" else ShardedAdvect)(*a, **kw)
