"""Host-side logic of the multi-GPU path, on CPU: point sharding, the row-band plan, and the counter
exchange (reduce-scatter by bands + halo all-gather) run for real over torch.distributed with the gloo
backend and world_size 2 and 3.  No CUDA here; the device kernels are covered by the gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clumpy_b200.multigpu import BandPlan, exchange_counts, field_rows, shard_indices, shard_range, strip_edges


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 16, 12392, 1 << 24):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("height,world", [(4096, 8), (16384, 8), (37, 2), (5, 8), (1000, 3)])
def test_field_rows_tile_the_image(height, world):
    spans = [field_rows(height, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == height
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.parametrize("how", ["home_rows", "index"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_indices_partition_the_points(how, world):
    rng = np.random.default_rng(5)
    h = 200
    clouds = {
        "uniform": (rng.random((50000, 2)) * [300, h]).astype(np.float32),
        "clustered": (rng.normal([150, 40], [30, 6], (50000, 2))).astype(np.float32),
        "one_row": np.full((1000, 2), 17.5, np.float32),
        "weird": np.array([[1, np.nan], [2, np.inf], [3, -np.inf], [4, -5], [5, 1e12], [6, 3.2]], np.float32),
        "empty": np.zeros((0, 2), np.float32),
    }
    for name, pts in clouds.items():
        parts = [shard_indices(pts, h, world, r, how) for r in range(world)]
        merged = np.concatenate(parts) if parts else np.zeros(0, int)
        assert sorted(merged.tolist()) == list(range(len(pts))), name
        for part in parts:
            assert (np.diff(part) > 0).all(), name  # ascending: a shard keeps the caller's relative order
        if how == "home_rows" and name in ("uniform", "clustered"):
            sizes = [len(p) for p in parts]
            per_row = np.bincount(np.clip(pts[:, 1], 0, h - 1).astype(int), minlength=h).max()
            assert max(sizes) - min(sizes) <= 2 * per_row + 1, (name, sizes)  # balanced to within a row of points
            # a strip is a contiguous range of home rows
            rows = [np.clip(pts[p, 1], 0, h - 1).astype(int) for p in parts if len(p)]
            assert all(a.max() < b.min() for a, b in zip(rows, rows[1:])), name


@pytest.mark.parametrize("height,halo,world", [(4096, 1, 8), (500, 0, 2), (500, 2, 3), (64, 3, 2), (2000, 3, 4), (37, 1, 2)])
def test_band_plan_partitions_the_image(height, halo, world):
    rows_alloc = -(-(-(-height // 16) * 16 + 2 * halo) // (16 * world)) * 16 * world
    plan = BandPlan(height, halo, rows_alloc, world)
    owned = [plan.owned_rows(r) for r in range(world)]
    covered = np.zeros(height, int)
    for r, (y0, y1) in enumerate(owned):
        covered[y0:y1] += 1
        lo, hi = plan.padded_rows(r)
        # every owned row's window of padded rows [y, y + 2*halo] lies inside chunk + halos
        if y1 > y0:
            assert y0 >= lo - halo and (y1 - 1) + 2 * halo <= hi + halo - 1
            assert plan.cell_row_offset(r) >= 0
            assert plan.cell_row_offset(r) + (-(-(y1 - y0) // 16) * 16) + 2 * halo <= plan.band_buffer_rows()
    assert (covered == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, height, halo, pitch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows_alloc = -(-(-(-height // 16) * 16 + 2 * halo) // (16 * world)) * 16 * world
    plan = BandPlan(height, halo, rows_alloc, world)
    images = [torch.from_numpy(np.random.default_rng(100 + r).integers(0, 9, (rows_alloc, pitch)).astype(np.float32))
              for r in range(world)]
    total = sum(images)
    band = torch.full((plan.band_buffer_rows(), pitch), -1.0)
    exchange_counts(images[rank], band, plan, rank, dist)
    lo, hi = plan.padded_rows(rank)
    ok = torch.equal(band[halo:halo + plan.chunk], total[lo:hi])
    if halo:
        above = total[lo - halo:lo] if rank > 0 else torch.zeros(halo, pitch)
        below = total[hi:hi + halo] if rank < world - 1 else torch.zeros(halo, pitch)
        ok = ok and torch.equal(band[:halo], above) and torch.equal(band[halo + plan.chunk:halo + plan.chunk + halo], below)
    # what the composite pass would read for the rows this rank owns equals the global sum there
    y0, y1 = plan.owned_rows(rank)
    off = plan.cell_row_offset(rank)
    if y1 > y0:
        ok = ok and torch.equal(band[off:off + (y1 - y0) + 2 * halo], total[y0:y1 + 2 * halo])
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,halo", [(2, 64, 1), (2, 100, 3), (3, 96, 2), (2, 48, 0)])
def test_exchange_counts_over_gloo(world, height, halo):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, height, halo, 40, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)), dict(out)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_strip_edges_agree_with_shard_indices(world):
    """The row strips a rank derives from the row histogram (what the device-side shard selection uses, and what the
    C++ driver restates in csrc/advect_multi.cu) are the strips shard_indices cuts on the host."""
    rng = np.random.default_rng(11)
    h = 137
    for pts in ((rng.random((20000, 2)) * [90, h]).astype(np.float32),
                rng.normal([40, 30], [10, 9], (20000, 2)).astype(np.float32),
                np.full((500, 2), 5.5, np.float32), np.zeros((0, 2), np.float32)):
        row = np.clip(np.nan_to_num(pts[:, 1].astype(np.float64)), 0, h - 1).astype(np.int64)
        rows = np.bincount(row, minlength=h)
        edges = strip_edges(rows, world)
        assert edges[0] == 0 and edges[-1] == h and all(a <= b for a, b in zip(edges, edges[1:]))
        for r in range(world):
            want = shard_indices(pts, h, world, r)
            got = np.nonzero((row >= edges[r]) & (row < edges[r + 1]))[0]
            assert np.array_equal(want, got)


def test_band_plan_ranks_without_rows():
    """More padded rows than the image has (8 ranks on 200 rows): the last rank's chunk lies in the padding and it
    owns no rows -- and says so with an empty range inside the image."""
    plan = BandPlan(200, 1, 256, 8)
    spans = [plan.owned_rows(r) for r in range(8)]
    assert spans[-1][0] == spans[-1][1] <= 200
    covered = sorted(y for a, b in spans for y in range(a, b))
    assert covered == list(range(200))
