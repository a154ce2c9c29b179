"""Pins the CPU oracle (oracle/clumpy_oracle.c) to outputs of the compiled reference.

CPU only.  The golden files were produced by tests/golden/make_golden.py from oracle/_ref/clumpy_ref,
i.e. the reference's own sources compiled with its own flags.  Integer / byte results must be
bit-identical; the float fields are bit-identical too because the restatement keeps the reference's
operation order and the build forbids FMA contraction.
"""
import hashlib

import os

import numpy as np
import pytest


@pytest.mark.parametrize("nframes", [1, 2, 7, 240, 1000, 65536, 1000003, 1500000000, 2**31 - 1, 2**30 + 7])
def test_age_offsets_equal_the_standard_library(oracle_mod, nframes):
    """The restated distribution (Lemire multiply-shift with rejection) against <random> itself, including
    ranges where a large share of the draws is rejected (every rejection shifts all later draws)."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref was not built here")
    n = 400000
    assert np.array_equal(oracle_mod.age_offsets(nframes, n), oracle_mod.ref_std_age_offsets(nframes, n))


def test_perm_and_age_offsets_known_answers(oracle_mod, golden):
    m = golden["meta"]
    assert oracle_mod.simplex_perm(0)[:8].tolist() == m["perm_seed0_first8"]
    p = oracle_mod.simplex_perm(12345)
    assert sorted(p.tolist()) == list(range(256))  # a permutation
    for nframes, first in m["age_offsets_first10"].items():
        assert oracle_mod.age_offsets(int(nframes), 10).tolist() == first
    o = oracle_mod.age_offsets(7, 100000)
    assert o.min() == 0 and o.max() == 6


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_simplex_and_curl_bit_identical(oracle_mod, golden, name):
    meta = golden["meta"][f"simplex_{name}"]
    dims, amp, freq, seed = meta["args"]
    w, h = (int(v) for v in dims.split("x"))
    pot, lo, hi = oracle_mod.generate_simplex(w, h, float(amp), float(freq), int(seed))
    ref = golden["fields"][f"pot_{name}"]
    assert pot.shape == ref.shape and np.array_equal(pot.view(np.uint32), ref.view(np.uint32))
    assert lo == ref.min() and hi == ref.max()
    vel, clo, chi = oracle_mod.curl_2d(ref)
    refv = golden["fields"][f"vel_{name}"]
    assert np.array_equal(vel.view(np.uint32), refv.view(np.uint32))
    # printed range keeps the reference's quirk: min over dpdx (channel 1) only
    assert clo == refv[..., 1].min() and chi == refv.max()
    stdout = golden["meta"][f"curl_{name}"]["stdout"][0]
    assert stdout == f"Curl range is {float(clo):g} to {float(chi):g}"


@pytest.mark.parametrize("name", ["k1", "k3", "k5", "k7", "k3_wrap", "k1_decay0", "k9"])
def test_advect_small_bit_identical(oracle_mod, golden, name):
    c = golden["meta"][f"advect_{name}"]
    adv = golden["advect"]
    frames, _ = oracle_mod.advect_points(adv["pts"], adv[c["vel"]], c["step"], c["k"], c["decay"], c["nframes"])
    assert np.array_equal(frames, adv[f"frames_{name}"])
    assert frames.any()


@pytest.mark.parametrize("name", ["k1_a1", "k3_a1", "k5_a05", "k1_a03", "k7_a1"])
def test_splat_command_bit_identical(oracle_mod, golden, name):
    c = golden["meta"][f"splat_{name}"]
    pts = golden["splat"]["pts"]
    if c["k"] == 1 and c["alpha"] == 1.0:  # splat_points.cc:93-94 fast path
        img = oracle_mod.splat_points(pts, 96, 64)
    else:
        img = oracle_mod.splat_disks(pts, 96, 64, c["alpha"], c["k"])
    assert np.array_equal(img, golden["splat"][name])


def test_single_hit_values(oracle_mod):
    """one hit on black: k=1 -> 127; k=3 -> centre 165, edge 127, corner 89."""
    one = np.array([[5.5, 5.5]], np.float32)
    assert oracle_mod.splat_disks(one, 11, 11, 1.0, 1)[5, 5] == 127
    img = oracle_mod.splat_disks(one, 11, 11, 1.0, 3)
    assert (img[5, 5], img[5, 4], img[4, 4]) == (165, 127, 89)
    assert np.count_nonzero(img) == 9
    with pytest.raises(ValueError):
        oracle_mod.splat_disks(one, 11, 11, 1.0, 4)


def test_c1_example5_hashes(oracle_mod, golden):
    """README example5 end to end on the CPU oracle against the reference's md5 hashes."""
    m = golden["meta"]
    pot, lo, hi = oracle_mod.generate_simplex(1000, 500, 1.0, 8.0, 0)
    assert m["c1_simplex_stdout"] == f"Noise range is {float(lo):g} to {float(hi):g}"
    vel, clo, chi = oracle_mod.curl_2d(pot)
    assert m["c1_curl_stdout"][0] == f"Curl range is {float(clo):g} to {float(chi):g}"
    frames, _ = oracle_mod.advect_points(golden["c1_pts"], vel, 30, 1, 0.95, 240)
    got = [hashlib.md5(f.tobytes()).hexdigest() for f in frames]
    assert got == m["c1_frame_data_md5"]
    assert int(np.count_nonzero(frames[0])) == m["c1_frame0"]["nonzero"] and int(frames[0].max()) == m["c1_frame0"]["max"]


def test_reset_schedule_closed_form(oracle_mod):
    """The age bookkeeping equals: point i is back at its origin after sim frame s iff
    s % nframes == offset_i (what the CUDA kernel implements)."""
    rng = np.random.default_rng(0)
    n, nframes = 500, 9
    pts = (rng.random((n, 2), np.float32) * [64, 48]).astype(np.float32)
    vel = np.full((48, 64, 2), 0.25, np.float32)  # constant drift: positions reveal the age
    a = oracle_mod.Advect(pts, vel, 1.0, 1, 0.5, nframes)
    off = oracle_mod.age_offsets(nframes, n)
    age = np.zeros(n, np.int64)
    started = np.zeros(n, bool)
    for s in range(2 * nframes):
        a.step()
        age += 1
        hit = (s % nframes) == off
        age[hit] = 0
        started |= hit
        cur = a.points()
        back = np.all(cur == pts, axis=1)
        assert np.array_equal(back, hit), s
    a.close()


@pytest.mark.skipif("not __import__('oracle').have_ref()")
def test_against_live_reference(oracle_mod, tmp_path):
    """Where oracle/_ref was built: random cases through the reference's own splat_disks and
    simplex entry points, including wrap mode and out-of-range / negative coordinates."""
    rng = np.random.default_rng(3)
    pts = ((rng.random((4000, 2)) - 0.1) * [140, 90]).astype(np.float32)
    pts[:5] = [[-3, -3], [1e12, 5], [np.nan, 4], [-1e12, 2], [127.999, 79.999]]
    for k, alpha, wrap in [(1, 1.0, False), (3, 1.0, False), (5, 0.7, True), (7, 1.0, True), (3, 0.2, False)]:
        if wrap:
            sel = np.isfinite(pts).all(axis=1) & (np.abs(pts) < 1e9).all(axis=1)
            p = pts[sel]  # signed overflow in the reference's wrap arithmetic is undefined behaviour
        else:
            p = pts
        base = rng.integers(0, 256, (80, 128), dtype=np.uint8)
        mine = oracle_mod.splat_disks(p, 128, 80, alpha, k, wrap, img=base)
        theirs = oracle_mod.ref_splat_disks(p, 128, 80, alpha, k, wrap, img=base)
        assert np.array_equal(mine, theirs), (k, alpha, wrap)
    xy = rng.random((20000, 2)) * 4000 - 2000
    for seed in (0, 77, -5):
        perm = oracle_mod.simplex_perm(seed)
        mine = np.array([oracle_mod.simplex2(perm, x, y) for x, y in xy[:3000]])
        theirs = oracle_mod.ref_simplex2(seed, xy[:3000])
        assert np.array_equal(mine, theirs)


# ---- the producers either side of the path: oracle vs the compiled reference ----------

def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["b0", "b1", "b2", "b3"])
def test_bridson_points_golden(oracle_mod, golden, name):
    dims, radius, seed = golden["meta"][f"bridson_{name}"]["args"]
    w, h = (int(v) for v in dims.split("x"))
    ref = golden["next"][f"bridson_{name}"]
    mine = oracle_mod.bridson_points(w, h, float(radius), int(seed))
    assert mine.shape == ref.shape
    assert golden["meta"][f"bridson_{name}"]["stdout"] == f"Generated {len(ref)} points."
    assert np.array_equal(_bits(mine), _bits(ref))


def test_bridson_points_c1_and_c2_lists(oracle_mod, golden):
    """The point lists of README example5 / example6 (bridson_points 1000x500 5 0 and 4000x2000 20 0)."""
    import hashlib
    c1 = oracle_mod.bridson_points(1000, 500, 5.0, 0)
    assert np.array_equal(_bits(c1), _bits(golden["c1_pts"])) and len(c1) == 12392
    c2 = oracle_mod.bridson_points(4000, 2000, 20.0, 0)
    assert hashlib.md5(c2.tobytes()).hexdigest() == golden["meta"]["c2_pts_data_md5"] and len(c2) == 12392


@pytest.mark.parametrize("name", ["p0", "p1", "p2"])
def test_pendulum_phase_golden(oracle_mod, golden, name):
    dims, friction, sx, sy = golden["meta"][f"pendulum_{name}"]["args"]
    w, h = (int(v) for v in dims.split("x"))
    mine = oracle_mod.pendulum_phase(w, h, float(friction), float(sx), float(sy))
    assert np.array_equal(_bits(mine), _bits(golden["next"][f"pendulum_{name}"]))


def test_pendulum_phase_c2_field(oracle_mod, golden):
    import hashlib
    field = oracle_mod.pendulum_phase(4000, 2000, 0.9, 2.0, 5.0)
    assert hashlib.md5(field.tobytes()).hexdigest() == golden["meta"]["c2_field_data_md5"]
    assert golden["meta"]["c2_md5"]["field.npy"] == "c6107b1fa21605d0a8c920f02b6e6710"


def test_cull_points_golden(oracle_mod, golden):
    nxt = golden["next"]
    mine = oracle_mod.cull_points(nxt["cull_pts"], nxt["cull_mask"])
    assert np.array_equal(_bits(mine), _bits(nxt["cull_out"]))
    assert golden["meta"]["cull"]["stdout"] == f"Removed {len(nxt['cull_pts']) - len(mine)} points."


@pytest.mark.skipif(not os.path.exists("/root/reference/commands/bridson_points.cc"), reason="needs the reference sources")
def test_producers_against_the_compiled_reference(oracle_mod, tmp_path):
    """Where the reference can be run (the build container): more shapes than the committed fixtures."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    run = lambda *a: oracle_mod.ref_cli(*a, cwd=str(tmp_path))
    for (w, h, r, seed) in [(300, 300, 0.8, 11), (128, 32, 3.3, 2), (90, 90, 6.0, -3)]:
        run("bridson_points", f"{w}x{h}", str(r), str(seed), "p.npy")
        assert np.array_equal(_bits(np.load(tmp_path / "p.npy")), _bits(oracle_mod.bridson_points(w, h, r, seed)))
    for (w, h, fr, sx, sy) in [(500, 500, 0.01, 1.0, 5.0), (77, 200, 1.5, 0.1, 9.0)]:
        run("pendulum_phase", f"{w}x{h}", str(fr), str(sx), str(sy), "f.npy")
        assert np.array_equal(_bits(np.load(tmp_path / "f.npy")), _bits(oracle_mod.pendulum_phase(w, h, fr, sx, sy)))


def test_simplex_rows_equal_the_whole_image_and_the_reference_corner(oracle_mod, golden):
    """The row-band entry of the restatement (used for checks at 16384 x 16384) equals the whole-image one, and
    the 4-octave sum of its rows equals the corner of the reference's own 16384 x 16384 octaves (summed in FP32)."""
    whole, _, _ = oracle_mod.generate_simplex(300, 200, 1.5, 12.0, 3)
    rows, _, _ = oracle_mod.generate_simplex_rows(300, 200, 37, 50, 1.5, 12.0, 3)
    assert np.array_equal(rows.view(np.uint32), whole[37:87].view(np.uint32))
    corner = golden["c3_corner"]
    total = None
    for amp, freq, seed in golden["meta"]["c3_octaves"]:
        o, _, _ = oracle_mod.generate_simplex_rows(16384, 16384, 0, corner.shape[0], amp, freq, int(seed))
        total = o if total is None else (total + o).astype(np.float32)
    assert np.array_equal(total[:, :corner.shape[1]].view(np.uint32), corner.view(np.uint32))
