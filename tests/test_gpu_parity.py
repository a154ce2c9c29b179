"""GPU parity tests: the CUDA path, called through the C ABI (clumpy_b200 -> libclumpy_cuda.so),
against the CPU oracle on the same inputs and against the golden outputs of the compiled reference.

Bars (BASELINE.json north_star):
  curl field            bit-exact
  noise field           <= 1e-5 absolute
  point trajectories    bit-exact here (bar: 1e-4 relative)
  u8 frames             within +-1 LSB on >= 99.9 % of pixels; bit-exact for k = 1
"""
import os

import numpy as np
import pytest

from conftest import frame_agreement

pytestmark = pytest.mark.gpu

NOISE_TOL = 1e-5
FRAME_FRACTION = 0.999


# ---- generate_simplex ----------------------------------------------------------------------

@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_simplex_golden(ctx, golden, name):
    dims, amp, freq, seed = golden["meta"][f"simplex_{name}"]["args"]
    w, h = (int(v) for v in dims.split("x"))
    out, lo, hi = ctx.generate_simplex(w, h, float(amp), float(freq), int(seed))
    ref = golden["fields"][f"pot_{name}"]
    tol = NOISE_TOL * max(1.0, float(amp))
    assert np.abs(out - ref).max() <= tol
    assert abs(lo - ref.min()) <= tol and abs(hi - ref.max()) <= tol
    assert lo == out.min() and hi == out.max()


@pytest.mark.parametrize("w,h,freq,seed", [(1000, 500, 8.0, 0), (333, 777, 64.0, 5), (2048, 1024, 4096.0, 9),
                                            (1, 1, 3.0, 1), (7, 3, 16.0, 2)])
def test_simplex_vs_oracle(ctx, oracle_mod, w, h, freq, seed):
    ref, rlo, rhi = oracle_mod.generate_simplex(w, h, 1.0, freq, seed)
    out, lo, hi = ctx.generate_simplex(w, h, 1.0, freq, seed)
    assert np.abs(out - ref).max() <= NOISE_TOL
    assert abs(lo - rlo) <= NOISE_TOL and abs(hi - rhi) <= NOISE_TOL


def test_simplex_row_bands_equal_whole(ctx):
    """Row-sharded generation (how the field is split across GPUs) is bit-identical to one call."""
    import torch
    w, h = 640, 200
    whole, _, _ = ctx.generate_simplex(w, h, 1.0, 8.0, 4)
    buf = torch.empty((h, w), dtype=torch.float32, device="cuda")
    for row0, nrows in [(0, 64), (64, 100), (164, 36)]:
        ctx.generate_simplex_dev(w, h, row0, nrows, 1.0, 8.0, 4, buf.data_ptr() + row0 * w * 4)
    ctx.sync()
    assert np.array_equal(buf.cpu().numpy(), whole)


# ---- curl_2d ---------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_curl_golden_bit_exact(ctx, golden, name):
    out, lo, hi = ctx.curl_2d(golden["fields"][f"pot_{name}"])
    ref = golden["fields"][f"vel_{name}"]
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert lo == ref[..., 1].min() and hi == ref.max()


@pytest.mark.parametrize("w,h", [(1000, 500), (514, 33), (257, 19), (2, 1), (1, 5), (4096, 64)])
def test_curl_vs_oracle_bit_exact(ctx, oracle_mod, w, h):
    rng = np.random.default_rng(w * 31 + h)
    src = rng.standard_normal((h, w)).astype(np.float32)
    ref, rlo, rhi = oracle_mod.curl_2d(src)
    out, lo, hi = ctx.curl_2d(src)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert (lo, hi) == (rlo, rhi)


def test_curl_bands_with_halo_row(ctx, oracle_mod):
    """Row bands with the one-row halo pointer reproduce the single-image result exactly."""
    import torch
    rng = np.random.default_rng(1)
    w, h = 512, 96
    src = rng.standard_normal((h, w)).astype(np.float32)
    ref, _, _ = oracle_mod.curl_2d(src)
    d_src = torch.from_numpy(src).cuda()
    d_dst = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    bands = [(0, 40), (40, 24), (64, 32)]
    for row0, nrows in bands:
        nxt = d_src.data_ptr() + (row0 + nrows) * w * 4 if row0 + nrows < h else None
        ctx.curl_2d_dev(d_src.data_ptr() + row0 * w * 4, w, nrows, nxt, d_dst.data_ptr() + row0 * w * 8)
    ctx.sync()
    assert np.array_equal(d_dst.cpu().numpy().view(np.uint32), ref.view(np.uint32))


# ---- splat_points / splat_disks --------------------------------------------------------------

@pytest.mark.parametrize("name", ["k1_a1", "k3_a1", "k5_a05", "k1_a03", "k7_a1"])
def test_splat_golden(ctx, golden, name):
    c = golden["meta"][f"splat_{name}"]
    pts = golden["splat"]["pts"]
    ref = golden["splat"][name]
    if c["k"] == 1 and c["alpha"] == 1.0:
        img = ctx.splat_points(pts, 96, 64)
        assert np.array_equal(img, ref)
        return
    img = ctx.splat_disks(pts, 96, 64, c["alpha"], c["k"])
    frac, worst = frame_agreement(img, ref)
    assert frac >= FRAME_FRACTION, (frac, worst)
    if c["k"] == 1:
        assert np.array_equal(img, ref)


@pytest.mark.parametrize("k,alpha,wrap,n,w,h", [
    (1, 1.0, False, 20000, 300, 200), (3, 1.0, False, 50000, 300, 200), (5, 1.0, True, 30000, 257, 131),
    (7, 1.0, False, 8000, 130, 70), (3, 1.0, True, 40000, 128, 64), (9, 1.0, False, 5000, 200, 100),
    (15, 0.8, True, 3000, 100, 90), (3, 1.0, False, 0, 64, 64), (1, 0.5, False, 100000, 33, 17)])
def test_splat_disks_vs_oracle(ctx, oracle_mod, k, alpha, wrap, n, w, h):
    """Dense lists (count cells, ring-order replay): the +-1 LSB bar; k = 1 is exact."""
    rng = np.random.default_rng(k * 1000 + n)
    pts = ((rng.random((n, 2)) * 1.2 - 0.1) * [w, h]).astype(np.float32)
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = oracle_mod.splat_disks(pts, w, h, alpha, k, wrap, img=base)
    img = ctx.splat_disks(pts, w, h, alpha, k, wrapx=wrap, img=base)
    frac, worst = frame_agreement(img, ref)
    assert frac >= FRAME_FRACTION, (frac, worst)
    if k == 1:
        assert np.array_equal(img, ref)  # a single opacity class replays exactly


@pytest.mark.parametrize("k,alpha,wrap,n,w,h", [
    (3, 1.0, False, 3000, 300, 200), (5, 1.0, True, 2000, 257, 131), (7, 1.0, False, 1000, 130, 70),
    (7, 0.3, True, 4000, 400, 300), (5, 0.6, False, 5000, 256, 256), (15, 1.0, False, 300, 100, 90)])
def test_splat_disks_sparse_is_bit_exact(ctx, oracle_mod, k, alpha, wrap, n, w, h):
    """Sparse lists (8 * n <= pixels) take the exact-order path.  Random points rarely share a texel,
    so all but a handful of pixels see only single-point cells and must match the reference exactly;
    the rest stay within 1 LSB."""
    rng = np.random.default_rng(k * 77 + n)
    pts = ((rng.random((n, 2)) * 1.1 - 0.05) * [w, h]).astype(np.float32)
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = oracle_mod.splat_disks(pts, w, h, alpha, k, wrap, img=base)
    img = ctx.splat_disks(pts, w, h, alpha, k, wrapx=wrap, img=base)
    d = np.abs(img.astype(int) - ref.astype(int))
    assert (d == 0).mean() >= 0.997 and (d <= 1).mean() >= 0.9999 and d.max() <= 2, \
        (float((d == 0).mean()), float((d <= 1).mean()), int(d.max()))


def test_splat_weak_opacity_dense_is_documented_limit(ctx, oracle_mod):
    """Outside the reference's configs: ~44 weak hits (alpha 0.4) per pixel in ONE splat.  Ring-order
    replay is not within 1 LSB everywhere here ("where the composite pass is approximate");
    the test pins the size of the deviation so that it cannot grow unnoticed."""
    rng = np.random.default_rng(3040)
    w, h, n = 128, 64, 40000
    pts = ((rng.random((n, 2)) * 1.2 - 0.1) * [w, h]).astype(np.float32)
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = oracle_mod.splat_disks(pts, w, h, 0.4, 3, True, img=base)
    img = ctx.splat_disks(pts, w, h, 0.4, 3, wrapx=True, img=base)
    frac, worst = frame_agreement(img, ref)
    assert frac >= 0.99 and worst <= 3, (frac, worst)


def test_splat_everything_on_one_pixel(ctx, oracle_mod):
    """Maximum contention: 200k points on one texel saturate exactly like the sequential loop."""
    pts = np.full((200000, 2), 20.3, np.float32)
    for k in (1, 3, 5):
        ref = oracle_mod.splat_disks(pts, 48, 40, 1.0, k)
        img = ctx.splat_disks(pts, 48, 40, 1.0, k)
        assert np.array_equal(img, ref), k


def test_splat_weird_coordinates(ctx, oracle_mod):
    pts = np.array([[np.nan, 3], [3, np.nan], [np.inf, 1], [-np.inf, 1], [1e30, 2], [-1e30, 2], [-0.9, -0.9],
                    [63.999, 31.999], [64.0, 32.0], [2147483648.0, 5], [-2147483648.0, 5], [4294967296.0, 7]], np.float32)
    for k in (1, 3):
        assert np.array_equal(ctx.splat_disks(pts, 64, 32, 1.0, k), oracle_mod.splat_disks(pts, 64, 32, 1.0, k))
    assert np.array_equal(ctx.splat_points(pts, 64, 32), oracle_mod.splat_points(pts, 64, 32))


def test_even_kernel_rejected(ctx):
    import clumpy_b200
    with pytest.raises(clumpy_b200.ClumpyCudaError) as e:
        ctx.splat_disks(np.zeros((1, 2), np.float32), 8, 8, 1.0, 4)
    assert e.value.code == clumpy_b200.ERR_INVALID and "odd integer" in str(e.value)


# ---- advect_points ----------------------------------------------------------------------------

def run_both(ctx, oracle_mod, pts, vel, step, k, decay, nframes, check_every=1):
    """Steps the oracle and the GPU side by side; returns the worst frame agreement and whether all
    trajectories were bit-identical at every checked frame."""
    a = oracle_mod.Advect(pts, vel, step, k, decay, nframes)
    g = ctx.advect(pts, vel, step, k, decay, nframes)
    worst_frac, worst_abs, traj_equal = 1.0, 0, True
    for s in range(2 * nframes):
        a.step()
        g.step()
        if s % check_every == 0 or s == 2 * nframes - 1:
            traj_equal &= np.array_equal(a.points().view(np.uint32), g.points().view(np.uint32))
            frac, worst = frame_agreement(g.image(), a.image())
            worst_frac, worst_abs = min(worst_frac, frac), max(worst_abs, worst)
    a.close()
    g.close()
    return worst_frac, worst_abs, traj_equal


@pytest.mark.parametrize("name", ["k1", "k3", "k5", "k7", "k3_wrap", "k1_decay0", "k9"])
def test_advect_golden(ctx, golden, name):
    """Whole command through clumpy_cuda_advect_points against the reference CLI's frames."""
    c = golden["meta"][f"advect_{name}"]
    adv = golden["advect"]
    frames = ctx.advect_points(adv["pts"], adv[c["vel"]], c["step"], c["k"], c["decay"], c["nframes"])
    ref = adv[f"frames_{name}"]
    assert frames.shape == ref.shape
    # 453 points on 96 x 64: the sparse, exact-order path.  Every pixel is within 1 LSB and all but a
    # handful are identical to the reference CLI's frames (the handful: a texel holding two points
    # while a third point with an index between theirs hits the same pixel).
    check_exact_order(frames, ref, c["k"])


def check_exact_order(frames, ref, k):
    for f in range(len(ref)):
        d = np.abs(frames[f].astype(int) - ref[f].astype(int))
        assert d.max() <= 1 and (d == 0).mean() >= 0.998, (f, float((d == 0).mean()), int(d.max()))
    if k == 1:
        assert np.array_equal(frames, ref)


@pytest.mark.parametrize("cells", ["f16", "u32", "exact"])
@pytest.mark.parametrize("name", ["k1", "k3", "k5", "k7", "k3_wrap", "k1_decay0"])
def test_advect_golden_every_cell_type(ctx, golden, monkeypatch, cells, name):
    """The same fixtures through each accumulator type (CLUMPY_CUDA_CELLS forces it; half cells exist for
    k <= 3 only, larger sprites asked for them get 32-bit cells)."""
    monkeypatch.setenv("CLUMPY_CUDA_CELLS", cells)
    c = golden["meta"][f"advect_{name}"]
    adv = golden["advect"]
    frames = ctx.advect_points(adv["pts"], adv[c["vel"]], c["step"], c["k"], c["decay"], c["nframes"])
    ref = adv[f"frames_{name}"]
    if cells == "exact" or c["k"] == 1:
        check_exact_order(frames, ref, c["k"])
        return
    # count-only cells replay ring by ring: +-1 LSB; the k = 7 fixture (weak 0.028 ring, decay 0.8,
    # 3.6 hits per pixel and frame -- a density the default selection routes to the exact-order
    # path) is the worst case we know: 99.76 % within 1 LSB, max 3 (tools/parity_report.py).
    bar = 0.997 if name == "k7" else FRAME_FRACTION
    for f in range(len(ref)):
        frac, worst = frame_agreement(frames[f], ref[f])
        assert frac >= bar and worst <= 3, (cells, name, f, frac, worst)


@pytest.mark.parametrize("k,decay,nframes,step", [(1, 0.95, 12, 30.0), (3, 0.99, 8, 60.0), (5, 0.9, 6, 45.0),
                                                  (7, 0.97, 5, 40.0), (3, -0.95, 8, 80.0), (1, 0.0, 5, 30.0)])
def test_advect_vs_oracle_stepwise(ctx, oracle_mod, golden, k, decay, nframes, step):
    vel = golden["fields"]["vel_a"] + np.array([-0.01, 0.004], np.float32)
    rng = np.random.default_rng(k + nframes)
    pts = ((rng.random((6000, 2)) * 1.1 - 0.05) * [96, 64]).astype(np.float32)
    frac, worst, traj = run_both(ctx, oracle_mod, pts, vel, step, k, decay, nframes)
    assert traj, "trajectories differ from the reference arithmetic"
    assert frac >= FRAME_FRACTION, (frac, worst)
    if k == 1:
        assert frac == 1.0 and worst == 0


def test_advect_c1_example5(ctx, oracle_mod, golden):
    """README example5 at full size: fields from the GPU kernels, frames against the reference md5s
    (k = 1, so the frames must be bit-identical)."""
    import hashlib
    pot, _, _ = ctx.generate_simplex(1000, 500, 1.0, 8.0, 0)
    ref_pot, _, _ = oracle_mod.generate_simplex(1000, 500, 1.0, 8.0, 0)
    assert np.abs(pot - ref_pot).max() <= NOISE_TOL
    vel, _, _ = ctx.curl_2d(ref_pot)  # parity run: same potential as the reference pipeline
    frames = ctx.advect_points(golden["c1_pts"], vel, 30, 1, 0.95, 240)
    got = [hashlib.md5(f.tobytes()).hexdigest() for f in frames]
    assert got == golden["meta"]["c1_frame_data_md5"]


def test_advect_dense_k3_long(ctx, oracle_mod):
    """Several hits per pixel per frame for many frames: the +-1 LSB bound must not drift."""
    rng = np.random.default_rng(11)
    w, h, n = 160, 96, 60000  # ~3.9 points per pixel
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 3.0, 2)
    vel, _, _ = oracle_mod.curl_2d(pot)
    pts = (rng.random((n, 2)) * [w, h]).astype(np.float32)
    frac, worst, traj = run_both(ctx, oracle_mod, pts, vel, 25.0, 3, 0.99, 40, check_every=5)
    assert traj
    assert frac >= FRAME_FRACTION and worst <= 1, (frac, worst)


@pytest.mark.parametrize("regroup", ["0", "1", "3"])
def test_advect_regrouping_keeps_identity(ctx, oracle_mod, monkeypatch, regroup):
    """Periodic regrouping of the point records by tile must not change anything observable:
    trajectories (in the caller's order) stay bit-identical, frames stay within the bar."""
    monkeypatch.setenv("CLUMPY_CUDA_REGROUP", regroup)  # (the environment switches still override the options)
    monkeypatch.setenv("CLUMPY_CUDA_CELLS", "f16")
    rng = np.random.default_rng(21)
    w, h, n = 320, 200, 70001
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 5.0, 8)
    vel, _, _ = oracle_mod.curl_2d(pot)
    pts = ((rng.random((n, 2)) * 1.04 - 0.02) * [w, h]).astype(np.float32)
    frac, worst, traj = run_both(ctx, oracle_mod, pts, vel, 35.0, 3, 0.98, 7, check_every=3)
    assert traj
    assert frac >= FRAME_FRACTION and worst <= 1, (frac, worst)


def test_advect_ragged_and_empty(ctx, oracle_mod, golden):
    vel = golden["fields"]["vel_c"]  # 37 x 50: width not a multiple of 4 (byte path)
    for n in (0, 1, 2, 3, 255, 257):
        rng = np.random.default_rng(n)
        pts = (rng.random((n, 2)) * [50, 37]).astype(np.float32)
        frac, worst, traj = run_both(ctx, oracle_mod, pts, vel, 3.0, 3, 0.9, 3)
        assert traj and frac >= FRAME_FRACTION, (n, frac, worst)


def test_advect_frame_async_pipeline(ctx, oracle_mod, golden):
    """frame_async + ping-pong framebuffers deliver the same frames as synchronous reads."""
    import torch
    adv = golden["advect"]
    c = golden["meta"]["advect_k3"]
    g = ctx.advect(adv["pts"], adv[c["vel"]], c["step"], c["k"], c["decay"], c["nframes"])
    g.step(c["nframes"])
    bufs = [torch.empty((64, 96), dtype=torch.uint8).pin_memory() for _ in range(c["nframes"])]
    for f in range(c["nframes"]):
        g.step()
        g.frame_async(bufs[f].data_ptr())
    g.wait_frames()
    g.close()
    ref = adv["frames_k3"]
    for f in range(c["nframes"]):
        frac, _ = frame_agreement(bufs[f].numpy(), ref[f])
        assert frac >= FRAME_FRACTION


def test_advect_linearity_of_counts_large(ctx):
    """Size-independent property at a larger size: with decay 0 and k = 1 every recorded frame is
    exactly {0, f(0), f(f(0)), ...} per pixel count, so frame pixels only take values from the replay
    sequence of alpha = 0.5, and pixels without a point are 0; total lit pixels <= npts."""
    rng = np.random.default_rng(0)
    w, h, n = 2048, 1024, 400000
    pts = (rng.random((n, 2)) * [w, h]).astype(np.float32)
    vel = np.zeros((h, w, 2), np.float32)  # points stay where they are
    g = ctx.advect(pts, vel, 1.0, 1, 0.0, 2)
    g.step(3)
    img = g.image()
    g.close()
    counts = np.zeros((h, w), np.int64)
    np.add.at(counts, (pts[:, 1].astype(np.int64), pts[:, 0].astype(np.int64)), 1)
    seq = [0]
    for _ in range(12):
        seq.append(int(np.float32(np.float32(0.5) * np.float32(seq[-1])) + np.float32(127.5)))
    lut = np.array(seq + [seq[-1]] * 64)
    assert np.array_equal(img, lut[np.minimum(counts, len(lut) - 1)].astype(np.uint8))


# ---- multi-GPU path, single rank (the N-rank run is tools/check_multigpu.py under torchrun) ----------

def test_sharded_advect_world_of_one(ctx, oracle_mod, golden):
    """The count / exchange / composite-rows / finish split used across GPUs, run with a process group
    of one rank: must agree with the oracle like the fused path does."""
    import socket
    import torch
    import torch.distributed as dist
    from clumpy_b200.multigpu import ShardedAdvect
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        rng = np.random.default_rng(9)
        vel = golden["fields"]["vel_a"] + np.array([-0.01, 0.004], np.float32)
        pts = ((rng.random((9000, 2)) * 1.1 - 0.05) * [96, 64]).astype(np.float32)
        for k, decay in [(3, 0.97), (1, 0.9), (5, 0.95)]:
            run = ShardedAdvect(ctx, pts, vel, 50.0, k, decay, 6, dist, torch.device("cuda", 0))
            ref = oracle_mod.Advect(pts, vel, 50.0, k, decay, 6)
            for s_ in range(12):
                run.step()
                ref.step()
                frac, worst = frame_agreement(run.gather_image(), ref.image())
                assert frac >= FRAME_FRACTION, (k, s_, frac, worst)
            assert np.array_equal(run.points().view(np.uint32), ref.points().view(np.uint32))
            run.close()
            ref.close()
    finally:
        dist.destroy_process_group()
        for v in ("CLUMPY_CUDA_CELLS", "CLUMPY_CUDA_OVERLAP", "CLUMPY_CUDA_CELL_ROWS_MULTIPLE"):
            os.environ.pop(v, None)


# ---- the producers either side of the path ------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("w,h,friction,sx,sy", [(64, 48, 0.3, 0.7, 1.3), (333, 17, 0.0, 3.0, 0.25), (500, 500, 0.01, 1.0, 5.0),
                                                (1, 1, 0.5, 1.0, 1.0), (4001, 3, 0.9, 2.0, 5.0), (2, 1000, 1.5, 0.1, 9.0)])
def test_pendulum_phase_bit_exact(ctx, oracle_mod, w, h, friction, sx, sy):
    got = ctx.pendulum_phase(w, h, friction, sx, sy)
    assert np.array_equal(got.view(np.uint32), oracle_mod.pendulum_phase(w, h, friction, sx, sy).view(np.uint32))


@pytest.mark.gpu
def test_pendulum_phase_c2_field_and_row_bands(ctx, golden):
    import hashlib
    import torch
    field = ctx.pendulum_phase(4000, 2000, 0.9, 2.0, 5.0)
    assert hashlib.md5(field.tobytes()).hexdigest() == golden["meta"]["c2_field_data_md5"]
    # rows are independent: bands written by the device entry point reproduce the whole field
    w, h = 640, 101
    whole = ctx.pendulum_phase(w, h, 0.2, 1.5, 2.5)
    dev = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    for r0, n in ((0, 40), (40, 1), (41, 60)):
        ctx.pendulum_phase_dev(w, h, r0, n, 0.2, 1.5, 2.5, dev[r0].data_ptr())
    ctx.sync()
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), whole.view(np.uint32))


@pytest.mark.gpu
def test_cull_points_golden_and_edge_cases(ctx, oracle_mod, golden):
    nxt = golden["next"]
    got = ctx.cull_points(nxt["cull_pts"], nxt["cull_mask"])
    assert got.shape == nxt["cull_out"].shape and np.array_equal(got.view(np.uint32), nxt["cull_out"].view(np.uint32))
    rng = np.random.default_rng(3)
    for (w, h, n) in [(1, 1, 10), (257, 33, 100001), (1000, 500, 1 << 20), (64, 64, 255), (64, 64, 256), (64, 64, 257)]:
        mask = rng.standard_normal((h, w)).astype(np.float32)
        pts = ((rng.random((n, 2)) * 1.3 - 0.15) * [w, h]).astype(np.float32)
        pts[:4] = [[np.nan, 0], [-0.5, 0.5], [4e9, 1], [-1e30, 2]]
        ref = oracle_mod.cull_points(pts, mask)
        got = ctx.cull_points(pts, mask)
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (w, h, n)
    # nothing kept / everything kept / no points at all
    ones = np.ones((8, 8), np.float32)
    inside = (rng.random((1000, 2)) * 8).astype(np.float32)
    assert len(ctx.cull_points(inside, -ones)) == 0
    assert np.array_equal(ctx.cull_points(inside, ones), inside)
    assert ctx.cull_points(np.zeros((0, 2), np.float32), ones).shape == (0, 2)


# ---- several frames per advect launch (clumpy_cuda_advect_step groups up to 4) --------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("cells", ["auto", "f16", "u32", "exact"])
@pytest.mark.parametrize("k,decay,nframes,overlap,regroup", [(3, 0.97, 7, 1, 3), (1, 0.9, 5, -1, 3), (5, -0.95, 6, 1, 3), (3, 0.0, 5, 1, 3),
                                                             (3, -0.97, 9, 1, -1), (1, 0.95, 8, 1, 8)])
def test_fused_launches_equal_frame_by_frame(ctx, oracle_mod, cells, k, decay, nframes, overlap, regroup):
    """step(n) in one call (groups of up to 8 frames per advect launch and per composite launch, three sets of
    counter images, composites on the side stream) must leave exactly the state that n calls of step(1) leave
    -- which the tests above pin to the oracle -- in every cell mode, across the warm-up / recording boundary
    (decay 0: splatting starts there), across regroups, in x-wrap mode, and it must agree with the oracle."""
    w, h, n = 192, 128, 30000
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 5.0, 4)
    vel, _, _ = oracle_mod.curl_2d(pot)
    pts = ((np.random.default_rng(8).random((n, 2)) * 1.06 - 0.03) * [w, h]).astype(np.float32)
    one = ctx.advect(pts, vel, 45.0, k, decay, nframes, cells=cells, overlap=overlap, regroup=regroup, fuse=1)
    many = ctx.advect(pts, vel, 45.0, k, decay, nframes, cells=cells, overlap=overlap, regroup=regroup)
    ref = oracle_mod.Advect(pts, vel, 45.0, k, decay, nframes)
    try:
        done = 0
        for chunk in (1, 4, 3, 8, 2, 4):  # groups are cut at regroups and at the warm-up / recording boundary
            chunk = min(chunk, 2 * nframes - done)
            if chunk == 0:
                break
            many.step(chunk)
            for _ in range(chunk):
                one.step(1)
                ref.step()
            done += chunk
            assert np.array_equal(many.points().view(np.uint32), one.points().view(np.uint32)), done
            assert np.array_equal(many.image(), one.image()), done
            assert np.array_equal(many.points().view(np.uint32), ref.points().view(np.uint32)), done
            if cells in ("auto", "u32", "exact") or k <= 3:  # (forced table replay with k = 5 is outside its regime)
                frac, worst = frame_agreement(many.image(), ref.image())
                assert frac >= 0.999, (done, frac, worst)
    finally:
        one.close()
        many.close()
        ref.close()


@pytest.mark.gpu
def test_fused_launches_with_frame_copies_in_flight(ctx, oracle_mod):
    """The frame read-back (frame_async) still sees every frame when the frames in between are
    advanced by fused launches."""
    w, h, n, nframes = 256, 96, 40000, 6
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 6.0, 2)
    vel, _, _ = oracle_mod.curl_2d(pot)
    pts = (np.random.default_rng(1).random((n, 2)) * [w, h]).astype(np.float32)
    adv = ctx.advect(pts, vel, 50.0, 3, 0.96, nframes, overlap=1)
    ref = oracle_mod.Advect(pts, vel, 50.0, 3, 0.96, nframes)
    bufs = [np.empty((h, w), np.uint8) for _ in range(4)]
    want = []
    try:
        for i, chunk in enumerate((4, 3, 4, 1)):
            adv.step(chunk)
            adv.frame_async(bufs[i])      # the copy is in flight while the next groups run
            for _ in range(chunk):
                ref.step()
            want.append(ref.image())
        adv.wait_frames()
        for i in range(4):
            frac, worst = frame_agreement(bufs[i], want[i])
            assert frac >= 0.999, (i, frac, worst)
    finally:
        adv.close()
        ref.close()


# ---- round 2: group composite (composite_group.cu), recording, device-resident set-up ---------------------------

def same_trajectories(a, b):
    """Bit-identical coordinates; a NaN coordinate stays NaN (the payload bits of a NaN that went through an
    addition are the hardware's choice: x86 keeps the input's, the GPU returns its canonical NaN)."""
    a32, b32 = a.view(np.uint32), b.view(np.uint32)
    return bool(np.all((a32 == b32) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,k,decay", [(128, 8, 3000, 3, 0.9), (127, 9, 3000, 3, 0.95), (130, 17, 5000, 1, 0.99), (4, 4, 200, 3, 0.5),
                                           (1, 1, 50, 3, 0.9), (1000, 37, 60000, 3, 0.99), (257, 64, 40000, 3, -0.97), (3, 70, 500, 3, -0.9),
                                           (1, 20, 300, 3, -0.9), (2, 20, 300, 3, -0.9)])
def test_group_composite_shapes(ctx, oracle_mod, w, h, n, k, decay):
    """Image sizes around the composite's 128 x 8 warp tiles and 4-pixel lanes (ragged right edge, byte path,
    one-pixel images, x-wrap with images narrower than the sprite), half cells forced, groups of 1 .. 8 frames."""
    rng = np.random.default_rng(w * 1000 + h)
    vel = (rng.standard_normal((h, w, 2)) * 0.02).astype(np.float32)
    pts = ((rng.random((n, 2)) * 1.1 - 0.05) * [w, h]).astype(np.float32)
    nframes = 9
    g = ctx.advect(pts, vel, 20.0, k, decay, nframes, cells="f16")
    a = oracle_mod.Advect(pts, vel, 20.0, k, decay, nframes)
    try:
        for chunk in (1, 8, 5, 4):
            g.step(chunk)
            for _ in range(chunk):
                a.step()
            assert np.array_equal(a.points().view(np.uint32), g.points().view(np.uint32))
            frac, worst = frame_agreement(g.image(), a.image())
            assert frac >= FRAME_FRACTION and worst <= 1, (chunk, frac, worst)
            if k == 1:
                assert worst == 0
    finally:
        a.close()
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cells,k", [("f16", 3), ("u32", 5), ("exact", 3), ("f16", 1)])
def test_record_delivers_every_frame(ctx, oracle_mod, cells, k):
    """clumpy_cuda_advect_record: every frame of fused groups lands in the caller's (pinned) buffer, identical
    to stepping one frame at a time and reading the framebuffer."""
    import torch
    w, h, n, nframes = 200, 72, 25000 if cells != "exact" else 600, 7
    pot, _, _ = oracle_mod.generate_simplex(w, h, 1.0, 6.0, 5)
    vel, _, _ = oracle_mod.curl_2d(pot)
    pts = (np.random.default_rng(3).random((n, 2)) * [w, h]).astype(np.float32)
    rec = ctx.advect(pts, vel, 50.0, k, 0.96, nframes, cells=cells)
    one = ctx.advect(pts, vel, 50.0, k, 0.96, nframes, cells=cells, fuse=1)
    total = 2 * nframes
    buf = torch.zeros((total, h, w), dtype=torch.uint8).pin_memory()
    try:
        rec.record(3, buf.data_ptr())
        rec.record(total - 3, buf.data_ptr() + 3 * h * w)
        rec.wait_frames()
        for f in range(total):
            one.step(1)
            assert np.array_equal(buf[f].numpy(), one.image()), f
        assert np.array_equal(rec.image(), one.image())
    finally:
        rec.close()
        one.close()


@pytest.mark.gpu
def test_device_resident_setup_and_shards(ctx, oracle_mod):
    """clumpy_cuda_advect_begin_ex with points, field and offsets already on the device (the simplex -> curl ->
    advect chain without a host round trip), and with row shards selected on the device: the union of the
    shards' trajectories is the unsharded run, bit for bit."""
    import torch
    import clumpy_b200
    from clumpy_b200.multigpu import strip_edges
    w, h, n, nframes = 320, 192, 50000, 5
    d_pot = torch.empty((h, w), dtype=torch.float32, device="cuda")
    d_vel = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
    ctx.generate_simplex_dev(w, h, 0, h, 1.0, 5.0, 3, d_pot.data_ptr())
    ctx.curl_2d_dev(d_pot.data_ptr(), w, h, None, d_vel.data_ptr())
    ctx.sync()
    vel = d_vel.cpu().numpy()
    pts = ((np.random.default_rng(4).random((n, 2)) * 1.04 - 0.02) * [w, h]).astype(np.float32)
    pts[:5] = [[np.nan, 3], [5, np.inf], [7, -np.inf], [9, -4], [11, 1e9]]
    d_pts = torch.from_numpy(pts).cuda()
    offs = clumpy_b200.age_offsets(nframes, n)
    d_offs = torch.from_numpy(offs.astype(np.int32)).cuda()
    ref = oracle_mod.Advect(pts, vel, 60.0, 3, 0.97, nframes)
    dev = ctx.advect(d_pts.data_ptr(), d_vel.data_ptr(), 60.0, 3, 0.97, nframes, age_offset=d_offs.data_ptr(),
                     npts=n, width=w, height=h)
    rows = ctx.row_histogram(d_pts.data_ptr(), h, npts=n)
    yy = np.clip(np.nan_to_num(pts[:, 1].astype(np.float64), nan=0.0, posinf=float(h), neginf=0.0), 0, h - 1).astype(np.int64)
    assert np.array_equal(rows, np.bincount(yy, minlength=h))
    edges = strip_edges(rows, 3)
    shards = [ctx.advect(d_pts.data_ptr(), d_vel.data_ptr(), 60.0, 3, 0.97, nframes, npts=n, width=w, height=h,
                         cells="count", shard_rows=(edges[r], edges[r + 1])) for r in range(3)]
    try:
        assert sum(s.n for s in shards) == n
        for _ in range(2):
            dev.step(nframes)
            for s in shards:
                s.step(nframes)
            for _ in range(nframes):
                ref.step()
            want = ref.points()
            assert same_trajectories(dev.points(), want)
            frac, worst = frame_agreement(dev.image(), ref.image())
            assert frac >= FRAME_FRACTION and worst <= 1
            got = np.empty_like(want)
            seen = np.zeros(n, bool)
            for s in shards:
                pos, idx = s.points_indexed()
                got[idx] = pos
                seen[idx] = True
            assert seen.all() and same_trajectories(got, want)
    finally:
        ref.close()
        dev.close()
        for s in shards:
            s.close()


@pytest.mark.gpu
def test_age_offset_sentinel_never_resets(ctx, oracle_mod):
    """Caller-supplied offsets >= nframes mean "never resets" (the reference's own sentinel, advect_points.cc:150)
    instead of being wrapped into 16 bits."""
    w, h, n, nframes = 96, 64, 4000, 6
    rng = np.random.default_rng(12)
    vel = np.full((h, w, 2), 0.01, np.float32)
    pts = (rng.random((n, 2)) * [w, h]).astype(np.float32)
    offs = rng.integers(0, nframes, n).astype(np.uint32)
    offs[::3] = nframes          # never
    offs[1::7] = 65536 + 2       # would alias phase 2 if truncated to 16 bits
    g = ctx.advect(pts, vel, 30.0, 1, 0.9, nframes, age_offset=offs)
    g.step(2 * nframes)
    got = g.points()
    g.close()
    want = pts.copy()
    step = np.float32(30.0) * np.float32(0.01)
    for s in range(2 * nframes):
        inside = (want[:, 0] >= 0) & (want[:, 0] < w) & (want[:, 1] >= 0) & (want[:, 1] < h)  # outside: velocity 0
        want[inside] = want[inside] + step
        reset = offs == (s % nframes)
        want[reset] = pts[reset]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("nframes,n", [(1000, 16_777_216), (1, 5000), (2, 70_001), (240, 1_000_003), (65535, 300_000),
                                        (1_000_003, 200_000), (3 << 29, 150_000), (3 << 29, 2_000_000), (0x7FFFFFFF, 100_000),
                                        (0x80000000, 100_000), (7, 34_500_000), (524_160, 524_160), (1000, 524_161)])
def test_age_offsets_drawn_on_the_device_match_the_host(ctx, nframes, n):
    """advect_points.cc:127-135 on the device (csrc/age_offsets_dev.cu) against the host restatement, which
    tests/test_abi.py pins to <random> itself.  3 << 29 rejects a quarter of the generator's words (2^32 mod range
    = 2^30), the other ranges between none and a handful: both the straight and the compacting path run.  The
    generator runs as up to 64 stretches from a table of states (16.7 M draws: 33 of them; 34.5 M: all 64 and the
    open-ended tail; 524 160 is exactly one stretch)."""
    import torch
    import clumpy_b200
    out = torch.full((n + 7,), -559038737, dtype=torch.int32, device="cuda")  # 0xDEADBEEF: the slack past n must stay untouched
    torch.cuda.synchronize()
    ctx.age_offsets_dev(nframes, n, out.data_ptr())
    ctx.sync()
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[:n], clumpy_b200.age_offsets(nframes, n))
    assert (got[n:] == 0xDEADBEEF).all()


# ---- parity at the sizes of the BASELINE configs ("never parity-checked at their sizes") --------------

@pytest.mark.gpu
def test_advect_c2_example6(ctx, golden, c2_frames):
    """README example6 at full size (README.md:119-121, extras/example6.py:16-30): pendulum_phase 4000x2000 field,
    12 392 bridson points, kernel 5, decay 0.99, 400 recorded frames.  A sparse cloud: the exact-order cells.
    Frames 0 / 199 / 399 against the reference CLI's (whose files have the md5s SURVEY 8c lists): every pixel
    within 1 LSB and at least 99.95 % of them identical (no frame is identical as a whole: a texel that holds two
    points while a third point with an index between theirs hits the same pixel is replayed in a slightly different
    order)."""
    import hashlib
    m = golden["meta"]
    assert m["c2_frame_md5"][0] == "0653b18c35d09cfb201d30f75290bdb7" and m["c2_frame_md5"][399] == "293c59068e2acafb378d9980cc4350aa"
    field = ctx.pendulum_phase(4000, 2000, 0.9, 2.0, 5.0)
    assert hashlib.md5(field.tobytes()).hexdigest() == m["c2_field_data_md5"]
    pts = golden["c2_pts"]
    assert len(pts) == 12392
    keep, seen = {}, []

    def on_frame(f, px):
        seen.append(f)
        if f in (0, 199, 399):
            keep[f] = px.copy()
        return 0

    ctx.advect_points(pts, field, 2.5, 5, 0.99, 400, on_frame=on_frame)
    assert seen == list(range(400))
    for f in (0, 199, 399):
        ref = c2_frames[f"frame_{f:03}"]
        d = np.abs(keep[f].astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1 and (d == 0).mean() >= 0.9995, (f, int(d.max()), float((d == 0).mean()))


@pytest.mark.gpu
def test_splat_c5_one_million_points(ctx, golden):
    """BASELINE configs[4], smallest size: 1 M uniform points on 16384 x 8192, k = 3 (9 111 185 non-zero
    pixels, maximum 249); the value histogram is the reference CLI's."""
    rng = np.random.default_rng(0)
    x = rng.random(1 << 20, dtype=np.float32) * np.float32(16384)
    y = rng.random(1 << 20, dtype=np.float32) * np.float32(8192)
    img = ctx.splat_disks(np.stack([x, y], axis=1), 16384, 8192, 1.0, 3)
    want = golden["meta"]["c5_k3_1M"]
    assert int(np.count_nonzero(img)) == want["nonzero"] == 9111185
    assert int(img.max()) == want["max"] == 249
    hist = np.bincount(img.ravel(), minlength=256)
    assert np.abs(hist - np.array(want["histogram"])).sum() <= 2e-4 * want["nonzero"]


@pytest.mark.gpu
def test_fields_c3_c4_sizes(ctx, golden):
    """generate_simplex + curl_2d at 8192 x 4096 (the C4 field) and 16384 x 16384 (C3): the ranges the reference
    prints, texels sampled from the reference's files (1e-5), the curl of the device potential
    bit-exact on sampled rows, and the 4-octave sum (README.md:28-36) against the reference's on a corner."""
    import torch
    m = golden["meta"]

    def printed(line, prefix):
        lo, hi = line.replace(prefix, "").split(" to ")
        return float(lo), float(hi)

    for w, h, key in ((8192, 4096, "c4"), (16384, 16384, "c3")):
        pot = torch.empty((h, w), dtype=torch.float32, device="cuda")
        vel = torch.empty((h, w, 2), dtype=torch.float32, device="cuda")
        lo, hi = ctx.generate_simplex_dev(w, h, 0, h, 1.0, 8.0, 0, pot.data_ptr(), want_range=True)
        rlo, rhi = printed(m[f"{key}_simplex_stdout"], "Noise range is ")
        assert abs(lo - rlo) <= 2e-6 and abs(hi - rhi) <= 2e-6, (key, lo, hi)
        clo, chi = ctx.curl_2d_dev(pot.data_ptr(), w, h, None, vel.data_ptr(), want_range=True)
        rclo, rchi = printed(m[f"{key}_curl_stdout"][0], "Curl range is ")
        assert abs(clo - rclo) <= 1e-5 and abs(chi - rchi) <= 1e-5 and abs(clo - rclo) <= 2e-3 * abs(rclo), (key, clo, chi)
        ctx.sync()
        for r in (0, 1, h // 2, h - 2, h - 1):  # curl_2d.cc:57-71 on the device potential, row by row
            p = pot[r].cpu().numpy()
            below = pot[min(r + 1, h - 1)].cpu().numpy()
            v = vel[r].cpu().numpy()
            dpdx = np.concatenate([p[1:] - p[:-1], [np.float32(0)]]).astype(np.float32)
            assert np.array_equal(v[:, 0].view(np.uint32), (p - below).astype(np.float32).view(np.uint32)), (key, r)
            assert np.array_equal(v[:, 1].view(np.uint32), dpdx.view(np.uint32)), (key, r)
        if key == "c3":
            s = m["c3_samples"]
            for i, r in enumerate(s["rows"]):
                got_p = pot[r].cpu().numpy()[s["cols"]]
                got_v = vel[r].cpu().numpy()[s["cols"]]
                assert np.abs(got_p - np.array(s["pot"][i], np.float32)).max() <= 1e-5
                assert np.abs(got_v - np.array(s["vel"][i], np.float32)).max() <= 1e-5
        del pot, vel
    # the multi-octave field: octaves evaluated and summed in FP32 in the README's order
    octs = m["c3_octaves"]
    corner = golden["c3_corner"]
    rows, cols = corner.shape
    out = torch.empty((rows, 16384), dtype=torch.float32, device="cuda")
    ctx.generate_simplex_octaves_dev(16384, 16384, 0, rows, [o[0] for o in octs], [o[1] for o in octs], [int(o[2]) for o in octs], out.data_ptr())
    ctx.sync()
    assert np.abs(out[:, :cols].cpu().numpy() - corner).max() <= 1e-5 * sum(o[0] for o in octs)
