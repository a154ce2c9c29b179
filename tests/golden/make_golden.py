#!/usr/bin/env python3
"""Regenerates tests/golden/* from the UNMODIFIED reference, compiled by oracle/Makefile into
oracle/_ref/clumpy_ref.  Run in the build container (needs /root/reference):

    make -C oracle && python tests/golden/make_golden.py

Everything written here is an OUTPUT OF THE REFERENCE CLI on inputs that are either produced by the
reference CLI itself (generate_simplex, curl_2d, bridson_points) or small synthetic .npy files that
this script also stores.  The reference has no result-pinning tests of its own, so
these files are what pins the oracle (tests/test_oracle.py) and, through it, the CUDA path.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def main():
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle"
    tmp = tempfile.mkdtemp(prefix="golden_")
    run = lambda *a: oracle.ref_cli(*a, cwd=tmp)
    load = lambda n: np.load(os.path.join(tmp, n))
    meta = {}

    # ---- C1: README example5 (README.md:97-111), hashes only + the point list ------------------
    out = run("generate_simplex", "1000x500", "1.0", "8.0", "0", "potential.npy")
    meta["c1_simplex_stdout"] = out.strip()
    out = run("curl_2d", "potential.npy", "velocity.npy")
    meta["c1_curl_stdout"] = out.strip().splitlines()
    run("bridson_points", "1000x500", "5", "0", "pts.npy")
    out = run("advect_points", "pts.npy", "velocity.npy", "30", "1", "0.95", "240", "anim.npy")
    meta["c1_advect_stdout_md5"] = hashlib.md5(out.encode()).hexdigest()  # 480 progress bars + the "Generated ..." line
    meta["c1_md5"] = {n: md5(os.path.join(tmp, n)) for n in ("potential.npy", "velocity.npy", "pts.npy")}
    meta["c1_frame_md5"] = [md5(os.path.join(tmp, f"{i:03}anim.npy")) for i in range(240)]
    meta["c1_frame_data_md5"] = [hashlib.md5(load(f"{i:03}anim.npy").tobytes()).hexdigest() for i in range(240)]
    shutil.copy(os.path.join(tmp, "pts.npy"), os.path.join(HERE, "c1_pts.npy"))
    f0 = load("000anim.npy")
    meta["c1_frame0"] = {"nonzero": int(np.count_nonzero(f0)), "max": int(f0.max())}

    # ---- small fields: simplex at several shapes / frequencies / seeds, and their curls --------
    fields = {}
    for name, (dims, amp, freq, seed) in {
        "a": ("96x64", "1.0", "4.0", "7"),
        "b": ("64x96", "2.5", "100.0", "3"),      # tall image, high frequency, amplitude != 1
        "c": ("50x37", "1.0", "1000.0", "-12"),   # odd sizes, very high frequency, negative seed
    }.items():
        out = run("generate_simplex", dims, amp, freq, seed, f"pot_{name}.npy")
        fields[f"pot_{name}"] = load(f"pot_{name}.npy")
        meta[f"simplex_{name}"] = {"args": [dims, amp, freq, seed], "stdout": out.strip()}
        out = run("curl_2d", f"pot_{name}.npy", f"vel_{name}.npy")
        fields[f"vel_{name}"] = load(f"vel_{name}.npy")
        meta[f"curl_{name}"] = {"stdout": out.strip().splitlines()}
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)

    # ---- small advect runs ---------------------------------------------------------------------
    run("bridson_points", "96x64", "3", "1", "pts_small.npy")
    pts = load("pts_small.npy")
    rng = np.random.default_rng(5)
    # a few hand-placed points on and beyond the borders, duplicates and a far-away one
    extra = np.array([[0.0, 0.0], [95.99, 63.99], [-0.5, 10.0], [96.0, 5.0], [-1.0, -1.0], [40.2, -0.2],
                      [40.2, 64.5], [1e9, 3.0], [50.5, 30.5], [50.5, 30.5], [50.7, 30.1]], np.float32)
    pts = np.concatenate([pts, extra]).astype(np.float32)
    np.save(os.path.join(tmp, "pts_small.npy"), pts)
    vel = fields["vel_a"].copy()
    drift = vel + np.array([-0.012, 0.007], np.float32)  # pushes points off the left / bottom edges
    np.save(os.path.join(tmp, "vel_small.npy"), vel)
    np.save(os.path.join(tmp, "vel_drift.npy"), drift)
    cases = {
        # name: (velocity file, step, k, decay, nframes)
        "k1": ("vel_small.npy", "40", "1", "0.9", "6"),
        "k3": ("vel_small.npy", "60", "3", "0.95", "5"),
        "k5": ("vel_drift.npy", "50", "5", "0.99", "4"),
        "k7": ("vel_drift.npy", "80", "7", "0.8", "3"),
        "k3_wrap": ("vel_drift.npy", "90", "3", "-0.9", "5"),  # negative decay = x-wrap mode
        "k1_decay0": ("vel_small.npy", "40", "1", "0", "4"),   # warm-up skips decay+splat
        "k9": ("vel_small.npy", "30", "9", "0.7", "2"),        # generic (non-unrolled) kernel size
    }
    adv = {"pts": pts, "vel_small": vel, "vel_drift": drift}
    for name, (vfile, step, k, decay, nframes) in cases.items():
        run("advect_points", "pts_small.npy", vfile, step, k, decay, nframes, f"_{name}.npy")
        adv[f"frames_{name}"] = np.stack([load(f"{i:03}_{name}.npy") for i in range(int(nframes))])
        meta[f"advect_{name}"] = {"vel": vfile[:-4], "step": float(step), "k": int(k), "decay": float(decay),
                                  "nframes": int(nframes)}
    np.savez_compressed(os.path.join(HERE, "advect_small.npz"), **adv)

    # ---- splat_points command --------------------------------------------------------------------
    spl = {"pts": pts}
    for name, (k, alpha) in {"k1_a1": ("1", "1.0"), "k3_a1": ("3", "1.0"), "k5_a05": ("5", "0.5"),
                             "k1_a03": ("1", "0.3"), "k7_a1": ("7", "1.0")}.items():
        out = run("splat_points", "pts_small.npy", "96x64", "u8disk", k, alpha, f"splat_{name}.npy")
        spl[name] = load(f"splat_{name}.npy")
        meta[f"splat_{name}"] = {"k": int(k), "alpha": float(alpha), "stdout": out.strip()}
    np.savez_compressed(os.path.join(HERE, "splat_small.npz"), **spl)

    # ---- the producers either side of the path: bridson_points, pendulum_phase, cull_points
    nxt = {}
    for name, (dims, radius, seed) in {"b0": ("200x100", "4", "3"), "b1": ("64x300", "2.5", "-7"),
                                       "b2": ("37x19", "1", "5"), "b3": ("500x250", "10", "987")}.items():
        out = run("bridson_points", dims, radius, seed, f"{name}.npy")
        nxt[f"bridson_{name}"] = load(f"{name}.npy")
        meta[f"bridson_{name}"] = {"args": [dims, radius, seed], "stdout": out.strip()}
    for name, (dims, friction, sx, sy) in {"p0": ("64x48", "0.3", "0.7", "1.3"), "p1": ("333x17", "0.0", "3.0", "0.25"),
                                           "p2": ("50x50", "0.01", "1.0", "5.0")}.items():
        run("pendulum_phase", dims, friction, sx, sy, f"{name}.npy")
        nxt[f"pendulum_{name}"] = load(f"{name}.npy")
        meta[f"pendulum_{name}"] = {"args": [dims, friction, sx, sy]}
    # C2 (README example6): the field and the point list by hash only (64 MB / 99 KB)
    run("pendulum_phase", "4000x2000", "0.9", "2", "5", "field.npy")
    run("bridson_points", "4000x2000", "20", "0", "pts_c2.npy")
    meta["c2_md5"] = {"field.npy": md5(os.path.join(tmp, "field.npy")), "pts.npy": md5(os.path.join(tmp, "pts_c2.npy"))}
    meta["c2_field_data_md5"] = hashlib.md5(load("field.npy").tobytes()).hexdigest()
    meta["c2_pts_data_md5"] = hashlib.md5(load("pts_c2.npy").tobytes()).hexdigest()
    # C2 advect (README.md:119-121, extras/example6.py:16-30): all 400 frames by hash (file and data), three of
    # them -- 4000 x 2000, mostly black -- stored compressed so that the CUDA path's "exact order" regime (k = 5,
    # sparse cloud: identical on all but a handful of pixels, never more than 1 LSB off) can be checked at size
    run("advect_points", "pts_c2.npy", "field.npy", "2.5", "5", "0.99", "400", "phase.npy")
    meta["c2_frame_md5"] = [md5(os.path.join(tmp, f"{i:03}phase.npy")) for i in range(400)]
    meta["c2_frame_data_md5"] = [hashlib.md5(load(f"{i:03}phase.npy").tobytes()).hexdigest() for i in range(400)]
    assert meta["c2_frame_md5"][0] == "0653b18c35d09cfb201d30f75290bdb7" and meta["c2_frame_md5"][399] == "293c59068e2acafb378d9980cc4350aa", \
        "C2 frames differ from SURVEY.md 8(c)"
    shutil.copy(os.path.join(tmp, "pts_c2.npy"), os.path.join(HERE, "c2_pts.npy"))
    np.savez_compressed(os.path.join(HERE, "c2_frames.npz"), **{f"frame_{i:03}": load(f"{i:03}phase.npy") for i in (0, 199, 399)})
    for i in range(400):
        os.unlink(os.path.join(tmp, f"{i:03}phase.npy"))
    # the argument-error messages of the four headline commands (advect_points.cc:65-68,84-104,
    # splat_points.cc:45-48,58-61,64-100, generate_simplex.cc:44-47, curl_2d.cc:30-45), from the reference itself
    import subprocess
    def fails(*a):
        p = subprocess.run([oracle.REF_CLI] + [str(x) for x in a], cwd=tmp, capture_output=True, text=True)
        return {"args": [str(x) for x in a], "rc": p.returncode, "stdout": p.stdout.strip()}
    np.save(os.path.join(tmp, "pts3.npy"), np.zeros((4, 3), np.float32))
    np.save(os.path.join(tmp, "pts64.npy"), np.zeros((4, 2), np.float64))
    np.save(os.path.join(tmp, "pts1d.npy"), np.zeros(8, np.float32))
    np.save(os.path.join(tmp, "vel2d.npy"), np.zeros((8, 8), np.float32))
    np.save(os.path.join(tmp, "vel64.npy"), np.zeros((8, 8, 2), np.float64))
    np.save(os.path.join(tmp, "ok_pts.npy"), np.ones((4, 2), np.float32))
    np.save(os.path.join(tmp, "ok_vel.npy"), np.zeros((8, 8, 2), np.float32))
    np.save(os.path.join(tmp, "img3d.npy"), np.zeros((4, 4, 2), np.float32))
    np.save(os.path.join(tmp, "img64.npy"), np.zeros((4, 4), np.float64))
    meta["cli_errors"] = [
        fails("advect_points", "a", "b"),
        fails("advect_points", "pts3.npy", "ok_vel.npy", "1", "1", "0.9", "2", "x.npy"),
        fails("advect_points", "pts64.npy", "ok_vel.npy", "1", "1", "0.9", "2", "x.npy"),
        fails("advect_points", "ok_pts.npy", "vel2d.npy", "1", "1", "0.9", "2", "x.npy"),
        fails("advect_points", "ok_pts.npy", "vel64.npy", "1", "1", "0.9", "2", "x.npy"),
        fails("advect_points", "ok_pts.npy", "ok_vel.npy", "1", "4", "0.9", "2", "x.npy"),
        fails("splat_points", "a"),
        fails("splat_points", "ok_pts.npy", "8x8", "u8disk", "2", "1.0", "x.npy"),
        fails("splat_points", "ok_pts.npy", "8x8", "fp32sphere", "3", "1.0", "x.npy"),
        fails("splat_points", "ok_pts.npy", "8x8", "bogus", "3", "1.0", "x.npy"),
        fails("splat_points", "pts1d.npy", "8x8", "u8disk", "3", "1.0", "x.npy"),
        fails("splat_points", "pts3.npy", "8x8", "u8disk", "3", "1.0", "x.npy"),
        fails("splat_points", "pts64.npy", "8x8", "u8disk", "3", "1.0", "x.npy"),
        fails("generate_simplex", "8x8"),
        fails("curl_2d", "a"),
        fails("curl_2d", "img3d.npy", "x.npy"),
        fails("curl_2d", "img64.npy", "x.npy"),
    ]
    # the printed ranges at the C4 / C3 field sizes and a few texels of each
    out = run("generate_simplex", "8192x4096", "1.0", "8.0", "0", "pot_c4.npy")
    meta["c4_simplex_stdout"] = out.strip()
    meta["c4_curl_stdout"] = run("curl_2d", "pot_c4.npy", "vel_c4.npy").strip().splitlines()
    os.unlink(os.path.join(tmp, "vel_c4.npy"))
    out = run("generate_simplex", "16384x16384", "1.0", "8.0", "0", "pot_c3.npy")
    meta["c3_simplex_stdout"] = out.strip()
    meta["c3_curl_stdout"] = run("curl_2d", "pot_c3.npy", "vel_c3.npy").strip().splitlines()
    pot3 = np.load(os.path.join(tmp, "pot_c3.npy"), mmap_mode="r")
    vel3 = np.load(os.path.join(tmp, "vel_c3.npy"), mmap_mode="r")
    samples = {"rows": [0, 1, 4097, 8192, 16383], "cols": [0, 5, 1000, 8191, 16383]}
    samples["pot"] = [[float(pot3[r, c]) for c in samples["cols"]] for r in samples["rows"]]
    samples["vel"] = [[[float(vel3[r, c, 0]), float(vel3[r, c, 1])] for c in samples["cols"]] for r in samples["rows"]]
    meta["c3_samples"] = samples
    # C3 as a multi-octave field (README.md:28-36, extras/test.py:18-21: separate commands, summed by numpy in
    # FP32): a 2048 x 24 corner of the 4-octave sum at 16384 x 16384 -- what a fused multi-octave kernel must match
    octs = [("1.0", "8.0", "0"), ("0.5", "16.0", "1"), ("0.25", "32.0", "2"), ("0.125", "64.0", "3")]
    total = np.array(pot3[:24, :2048])
    for amp, freq, seed in octs[1:]:
        run("generate_simplex", "16384x16384", amp, freq, seed, "oct.npy")
        total = total + np.load(os.path.join(tmp, "oct.npy"), mmap_mode="r")[:24, :2048]
    assert total.dtype == np.float32
    meta["c3_octaves"] = [[float(a), float(f), int(s)] for a, f, s in octs]
    del pot3, vel3
    # C5: splat_points sweep, smallest size (1 M uniform points on 16384 x 8192, k = 3): SURVEY.md 8(c)
    rng5 = np.random.default_rng(0)
    x = rng5.random(1 << 20, dtype=np.float32) * np.float32(16384)
    y = rng5.random(1 << 20, dtype=np.float32) * np.float32(8192)
    np.save(os.path.join(tmp, "pts_c5.npy"), np.stack([x, y], axis=1))
    run("splat_points", "pts_c5.npy", "16384x8192", "u8disk", "3", "1.0", "splat_c5.npy")
    img5 = load("splat_c5.npy")
    meta["c5_k3_1M"] = {"nonzero": int(np.count_nonzero(img5)), "max": int(img5.max()),
                        "histogram": np.bincount(img5.ravel(), minlength=256).tolist()}
    assert meta["c5_k3_1M"]["nonzero"] == 9111185 and meta["c5_k3_1M"]["max"] == 249, "C5 differs from SURVEY.md 8(c)"
    rng = np.random.default_rng(1)
    mask = rng.standard_normal((60, 100)).astype(np.float32)
    cpts = ((rng.random((5000, 2)) * 1.2 - 0.1) * [100, 60]).astype(np.float32)
    cpts[:5] = [[np.nan, 1], [1e12, 2], [-1e12, 3], [np.inf, 1], [3.999, 4.999]]
    np.save(os.path.join(tmp, "mask.npy"), mask)
    np.save(os.path.join(tmp, "cpts.npy"), cpts)
    out = run("cull_points", "cpts.npy", "mask.npy", "culled.npy")
    nxt["cull_mask"], nxt["cull_pts"], nxt["cull_out"] = mask, cpts, load("culled.npy")
    meta["cull"] = {"stdout": out.strip()}
    np.savez_compressed(os.path.join(HERE, "next_small.npz"), **nxt)
    np.savez_compressed(os.path.join(HERE, "c3_octaves_corner.npz"), total=total)

    # ---- scalar goldens SURVEY.md 8(c) lists ---------------------------------------------------
    meta["age_offsets_first10"] = {"240": [131, 142, 171, 202, 144, 205, 130, 203, 101, 149],
                                   "400": [219, 237, 286, 337, 241, 343, 217, 338, 169, 249],
                                   "1000": [548, 592, 715, 844, 602, 857, 544, 847, 423, 623]}
    meta["perm_seed0_first8"] = [254, 50, 92, 24, 36, 10, 190, 16]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    shutil.rmtree(tmp)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
