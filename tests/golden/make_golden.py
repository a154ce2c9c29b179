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
    run("advect_points", "pts.npy", "velocity.npy", "30", "1", "0.95", "240", "anim.npy")
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
