"""The multi-GPU advect path (exchange fused into the advect kernel, device-side frame fence) run as
SEVERAL RANKS ON ONE GPU: every rank is a libclumpy_cuda context of its own (own streams) on cuda:0 and
the "peer" inboxes are plain device pointers, so the whole protocol -- home-row sharding, own-band
counting, per-CTA inbox regions, epoch flags, drain, band composite on the side stream, band buffers
cycling -- runs for real and is compared with the CPU oracle frame by frame.  The same code over
CUDA IPC / NVLink is checked by tools/check_multigpu.py under torchrun on a multi-GPU box.

Reference semantics: commands/advect_points.cc:119-179 and splat_points.cc:118-161 (via oracle/)."""
import os

import numpy as np
import pytest

from conftest import frame_agreement

pytestmark = pytest.mark.gpu


class RanksOnOneGpu:
    def __init__(self, world, pts, vel, step, k, decay, nframes, overlap, shard="home_rows", on_device=False):
        import clumpy_b200
        from clumpy_b200.multigpu import shard_indices, strip_edges
        self.world, self.n, self.h, self.w = world, len(pts), vel.shape[0], vel.shape[1]
        self.ctxs = [clumpy_b200.Context(0) for _ in range(world)]
        opts = dict(cells="count", overlap=1 if overlap else -1)
        if on_device:
            # the shards are selected on the device from the whole list (desc.shard_row_lo / _hi), offsets drawn
            # by the library: what bench.py and the CLI's multi-GPU driver do
            rows = self.ctxs[0].row_histogram(pts, self.h)
            edges = strip_edges(rows, world)
            self.advs = [c.advect(pts, vel, step, k, decay, nframes, shard_rows=(edges[r], edges[r + 1]) if edges[r + 1] > edges[r] else (self.h, self.h + 1), **opts)
                         for r, c in enumerate(self.ctxs)]
            self.idx = None
            sizes = [a.n for a in self.advs]
            assert sizes == [int(rows[edges[r]:edges[r + 1]].sum()) for r in range(world)]
        else:
            offsets = clumpy_b200.age_offsets(nframes, len(pts))
            self.idx = [shard_indices(pts, self.h, world, r, shard) for r in range(world)]
            self.advs = [c.advect(np.ascontiguousarray(pts[i]), vel, step, k, decay, nframes,
                                  age_offset=np.ascontiguousarray(offsets[i]), **opts)
                         for c, i in zip(self.ctxs, self.idx)]
            sizes = [len(i) for i in self.idx]
        capacity = max(1, max(sizes))
        inboxes = [a.route_begin(r, world, capacity)[0] for r, a in enumerate(self.advs)]
        for a in self.advs:
            a.route_peers(inboxes)
        self.rows = [a.owned_rows() for a in self.advs]

    def step(self, count=1):
        # every rank enqueues the same `count` frames (one route launch covers up to 4 of them); a
        # rank's drains wait (bounded) for the epochs of the ranks enqueued after it
        for a in self.advs:
            a.route_steps(count)

    def image(self):
        img = np.zeros((self.h, self.w), np.uint8)
        for a, (y0, y1) in zip(self.advs, self.rows):
            img[y0:y1] = a.image()[y0:y1]
        return img

    def points(self):
        out = np.empty((self.n, 2), np.float32)
        for r, a in enumerate(self.advs):
            if self.idx is None:
                pos, idx = a.points_indexed()
                out[idx] = pos
            else:
                out[self.idx[r]] = a.points()
        return out

    def close(self):
        errors = [a.route_errors() for a in self.advs]
        for a in self.advs:
            a.close()
        for c in self.ctxs:
            c.close()
        assert errors == [0] * self.world, f"frame fences timed out: {errors}"


def _field(oracle_mod, w, h, seed=8):
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 5.0, seed)
    vel, _, _ = oracle_mod.curl_2d(pot)
    return vel


@pytest.mark.parametrize("overlap,chunk,on_device", [(False, 1, False), (True, 1, True), (True, 8, True), (False, 3, False), (True, 5, False)])
@pytest.mark.parametrize("world,w,h,n,k,decay,nframes,step", [
    (2, 320, 200, 70001, 3, 0.98, 6, 35.0),
    (3, 256, 160, 30000, 1, 0.9, 5, 40.0),
    (4, 200, 136, 20000, 5, 0.95, 4, 30.0),
    (2, 512, 64, 50000, 3, -0.97, 4, 60.0),     # negative decay: x-wrap mode
    (8, 300, 520, 90000, 3, 0.97, 5, 90.0),     # eight bands, fast field: many points change band
    (2, 128, 96, 5000, 3, 0.0, 3, 25.0),        # decay 0: warm-up frames do not splat at all
])
def test_ranks_on_one_gpu_match_oracle(oracle_mod, overlap, chunk, on_device, world, w, h, n, k, decay, nframes, step):
    vel = _field(oracle_mod, w, h)
    # a cloud that also has points outside the image on every side
    pts = ((np.random.default_rng(21).random((n, 2)) * 1.04 - 0.02) * [w, h]).astype(np.float32)
    run = RanksOnOneGpu(world, pts, vel, step, k, decay, nframes, overlap, on_device=on_device)
    ref = oracle_mod.Advect(pts, vel, step, k, decay, nframes)
    try:
        for s in range(0, 2 * nframes, chunk):
            todo = min(chunk, 2 * nframes - s)  # groups of frames: fused route launches, crossing the warm-up boundary
            run.step(todo)
            for _ in range(todo):
                ref.step()
            assert np.array_equal(run.points().view(np.uint32), ref.points().view(np.uint32)), f"trajectories differ at frame {s}"
            frac, worst = frame_agreement(run.image(), ref.image())
            assert frac >= 0.999, f"frame {s}: only {frac:.5f} of the pixels within 1 LSB (max {worst})"
            if k == 1:
                assert worst == 0, f"frame {s}: k = 1 frames must be bit-identical"
    finally:
        ref.close()
        run.close()


def test_index_sharding_routes_everything_through_inboxes(oracle_mod):
    """Contiguous index ranges: a rank's points are all over the image, so (world-1)/world of them
    travel through inboxes every frame."""
    w, h, n, k, nframes = 256, 192, 60000, 3, 4
    vel = _field(oracle_mod, w, h, seed=3)
    pts = (np.random.default_rng(5).random((n, 2)) * [w, h]).astype(np.float32)
    run = RanksOnOneGpu(4, pts, vel, 50.0, k, 0.96, nframes, True, shard="index")
    ref = oracle_mod.Advect(pts, vel, 50.0, k, 0.96, nframes)
    try:
        for s in range(2 * nframes):
            run.step()
            ref.step()
        assert np.array_equal(run.points().view(np.uint32), ref.points().view(np.uint32))
        frac, worst = frame_agreement(run.image(), ref.image())
        assert frac >= 0.999, (frac, worst)
    finally:
        ref.close()
        run.close()


def test_all_points_in_one_band_and_an_empty_rank(oracle_mod):
    """Every point starts in the top rows: home-row sharding still balances the shards (strip edges are
    row quantiles), bands 1.. receive nothing at first, and a rank whose shard is empty still takes
    part in the fence."""
    w, h, nframes = 200, 160, 4
    vel = _field(oracle_mod, w, h, seed=11)
    rng = np.random.default_rng(9)
    pts = np.stack([rng.random(20000) * w, rng.random(20000) * 12.0], axis=1).astype(np.float32)
    for world, cloud in ((4, pts), (4, pts[:2])):  # 2 points on 4 ranks: two shards are empty
        run = RanksOnOneGpu(world, cloud, vel, 45.0, 3, 0.95, nframes, False, on_device=len(cloud) > 2)
        ref = oracle_mod.Advect(cloud, vel, 45.0, 3, 0.95, nframes)
        try:
            for s in range(2 * nframes):
                run.step()
                ref.step()
                frac, worst = frame_agreement(run.image(), ref.image())
                assert frac >= 0.999, (s, frac, worst)
            assert np.array_equal(run.points().view(np.uint32), ref.points().view(np.uint32))
        finally:
            ref.close()
            run.close()


def test_dead_peer_times_out_instead_of_hanging(oracle_mod, monkeypatch):
    """A rank whose peer never publishes its epoch must not hang the GPU: the drain gives up after its
    bounded wait (CLUMPY_ROUTE_TIMEOUT_MS), and the time-out is STICKY: every later call on the run fails with
    CLUMPY_ERR_STATE instead of handing out incomplete frames."""
    import clumpy_b200
    monkeypatch.setenv("CLUMPY_ROUTE_TIMEOUT_MS", "300")
    w, h = 128, 96
    vel = _field(oracle_mod, w, h)
    pts = (np.random.default_rng(2).random((4000, 2)) * [w, h]).astype(np.float32)
    run = RanksOnOneGpu(2, pts, vel, 20.0, 3, 0.9, 3, False)
    try:
        run.advs[0].route_steps(1)      # rank 1 never runs this frame
        assert run.advs[0].route_errors() >= 1
        for call in (lambda: run.advs[0].route_steps(1), lambda: run.advs[0].image()):
            with pytest.raises(clumpy_b200.ClumpyCudaError) as e:
                call()
            assert e.value.code == clumpy_b200.ERR_STATE and "timed out" in str(e.value)
    finally:
        for a in run.advs:
            a.close()
        for c in run.ctxs:
            c.close()


# ---- the C++ driver behind `clumpy advect_points` with CLUMPY_DEVICES: several ranks from one process ---------------

@pytest.mark.parametrize("devices,w,h,n,k,decay,nframes,step", [
    ([0, 0], 320, 200, 60000, 3, 0.97, 9, 40.0),
    ([0, 0, 0], 256, 160, 30000, 1, 0.9, 70, 40.0),     # more frames than one turn of the warm-up loop, regroups inside
    ([0, 0, 0, 0], 200, 136, 9000, 5, -0.95, 5, 30.0),  # 32-bit cells, x-wrap
    ([0], 128, 96, 20000, 3, 0.95, 4, 25.0),            # a world of one
])
def test_single_process_driver_matches_oracle(oracle_mod, devices, w, h, n, k, decay, nframes, step):
    """clumpy_cuda_advect_points_multi (advect_points.cc:137-179 on several GPUs from one process; here the "GPUs"
    are the same device listed several times): every recorded frame against the oracle's."""
    import clumpy_b200
    vel = _field(oracle_mod, w, h)
    pts = ((np.random.default_rng(33).random((n, 2)) * 1.04 - 0.02) * [w, h]).astype(np.float32)
    frames = clumpy_b200.advect_points_multi(devices, pts, vel, step, k, decay, nframes)
    ref = oracle_mod.Advect(pts, vel, step, k, decay, nframes)
    try:
        for _ in range(nframes):
            ref.step()
        for f in range(nframes):
            ref.step()
            frac, worst = frame_agreement(frames[f], ref.image())
            assert frac >= 0.999, (f, frac, worst)
            if k == 1:
                assert worst == 0, f
    finally:
        ref.close()


def test_cli_advect_points_on_several_devices(golden, tmp_path):
    """`CLUMPY_DEVICES=0,0 clumpy advect_points ...` (README example5, k = 1): the same 240 files, byte for byte, as
    the reference CLI and as the one-GPU run."""
    import hashlib
    import subprocess
    import oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pot, _, _ = oracle.generate_simplex(1000, 500, 1.0, 8.0, 0)
    vel, _, _ = oracle.curl_2d(pot)
    np.save(tmp_path / "velocity.npy", vel)
    np.save(tmp_path / "pts.npy", golden["c1_pts"])
    # CLUMPY_TEST_DEVICES=0,1 on a multi-GPU box: the same run over NVLink peer memory
    env = dict(os.environ, CLUMPY_DEVICES=os.environ.get("CLUMPY_TEST_DEVICES", "0,0"))
    p = subprocess.run([os.path.join(root, "clumpy_b200", "clumpy"), "advect_points", "pts.npy", "velocity.npy", "30", "1", "0.95", "240", "anim.npy"],
                       cwd=tmp_path, env=env, capture_output=True)
    assert p.returncode == 0, p.stdout.decode(errors="replace")[-300:]
    assert hashlib.md5(p.stdout).hexdigest() == golden["meta"]["c1_advect_stdout_md5"]
    got = [hashlib.md5(open(tmp_path / f"{i:03}anim.npy", "rb").read()).hexdigest() for i in range(240)]
    assert got == golden["meta"]["c1_frame_md5"]


# ---- splat_points on several devices (last row) ----------------------------------------------------------

def _test_devices(default):
    env = os.environ.get("CLUMPY_TEST_DEVICES")
    return [int(v) for v in env.split(",")] if env else default


@pytest.mark.parametrize("k,alpha,wrap,n,w,h", [
    (1, 1.0, False, 30000, 300, 200),     # the fast path of the command (stores 255)
    (3, 1.0, False, 50000, 300, 201),     # dense: 32-bit counts summed across the devices
    (5, 0.6, True, 30000, 257, 131),      # x-wrap
    (3, 1.0, False, 2000, 300, 200),      # sparse: exact-order cells (count + sum of ORIGINAL indices)
    (7, 0.3, True, 900, 400, 300),
    (1, 0.5, False, 5000, 33, 17),        # fewer rows than some band splits like
    (3, 1.0, False, 0, 64, 64),
])
def test_splat_multi_equals_one_device(ctx, k, alpha, wrap, n, w, h):
    """clumpy_cuda_splat_multi: the accumulators of the devices sum to the one-device accumulator, so the image must
    be IDENTICAL to clumpy_cuda_splat_points / _splat_disks (which the parity tests pin to the reference)."""
    import clumpy_b200
    rng = np.random.default_rng(k * 100 + n)
    pts = ((rng.random((n, 2)) * 1.1 - 0.05) * [w, h]).astype(np.float32)
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    one = ctx.splat_points(pts, w, h, img=base) if (k == 1 and alpha == 1.0) else ctx.splat_disks(pts, w, h, alpha, k, wrapx=wrap, img=base)
    for devices in (_test_devices([0, 0]), _test_devices([0, 0, 0]) + [0]):
        got = clumpy_b200.splat_multi(devices, pts, w, h, alpha, k, wrapx=wrap, img=base)
        assert np.array_equal(got, one), (devices, int(np.abs(got.astype(int) - one.astype(int)).max()))


def test_cli_splat_points_on_several_devices(golden, tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spl, m = golden["splat"], golden["meta"]
    np.save(tmp_path / "pts.npy", spl["pts"])
    env = dict(os.environ, CLUMPY_DEVICES=os.environ.get("CLUMPY_TEST_DEVICES", "0,0"))
    for name in ("k1_a1", "k3_a1", "k7_a1"):
        c = m[f"splat_{name}"]
        p = subprocess.run([os.path.join(root, "clumpy_b200", "clumpy"), "splat_points", "pts.npy", "96x64", "u8disk", str(c["k"]), str(c["alpha"]), f"{name}.npy"],
                           cwd=tmp_path, env=env, capture_output=True)
        assert p.returncode == 0 and p.stdout.decode().strip() == c["stdout"]
        d = np.abs(np.load(tmp_path / f"{name}.npy").astype(int) - spl[name].astype(int))
        assert d.max() <= 1 and (d <= 1).mean() >= 0.999, name
        if c["k"] == 1:
            assert d.max() == 0
