"""The host CLI (clumpy_b200/clumpy): same command names, arguments, messages, exit codes and .npy
bytes as the reference's `clumpy` (main_clumpy.cc:37-60, commands/*.cc).  CPU tests cover what needs
no device (dispatch, argument checks, bridson_points -- which is a host command by design); gpu tests
run the device-backed producers through the CLI and compare FILE BYTES with the compiled reference's
(tests/golden/golden.json holds the md5 of its output files)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "clumpy_b200", "clumpy")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-s", "-C", ROOT, "-j8", "lib"], check=True)
        subprocess.run(["make", "-s", "-C", ROOT, "cli"], check=True)

    def run(*args, cwd=None):
        p = subprocess.run([CLI] + [str(a) for a in args], cwd=cwd, capture_output=True)
        return p.returncode, p.stdout.decode(errors="replace")  # (bytes as printed: no newline translation of the \r's)
    return run


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_help_lists_every_command_of_the_path(cli):
    rc, out = cli("help")
    assert rc == 0
    for name in ("generate_simplex", "curl_2d", "advect_points", "splat_points", "bridson_points",
                 "pendulum_phase", "cull_points"):
        assert f"\n{name} " in out, name
    rc, out = cli("no_such_command")
    assert rc == 1 and "help" in out  # main_clumpy.cc:48-51: unknown command -> help text, exit 1


@pytest.mark.parametrize("cmd,nargs,msg", [
    ("bridson_points", 2, "This command takes 4 arguments."),
    ("pendulum_phase", 6, "The command takes 5 arguments."),
    ("cull_points", 1, "This command takes 3 arguments."),
])
def test_wrong_argument_count_prints_the_reference_message(cli, cmd, nargs, msg):
    rc, out = cli(cmd, *["x"] * nargs)
    assert rc == 1 and out.strip() == msg


def test_bridson_points_files_are_byte_identical_to_the_reference(cli, golden, tmp_path):
    """README example5 / example6 point lists (bridson_points.cc:38-177)."""
    rc, out = cli("bridson_points", "1000x500", "5", "0", "pts.npy", cwd=tmp_path)
    assert rc == 0 and out.strip() == "Generated 12392 points."
    assert md5(tmp_path / "pts.npy") == golden["meta"]["c1_md5"]["pts.npy"] == "2d999229f060611feff0d05be8318b13"
    rc, out = cli("bridson_points", "4000x2000", "20", "0", "pts2.npy", cwd=tmp_path)
    assert rc == 0 and md5(tmp_path / "pts2.npy") == golden["meta"]["c2_md5"]["pts.npy"] == "ca686837e37fbe6a04f3c767af444c25"
    for name in ("b0", "b1", "b2", "b3"):
        dims, radius, seed = golden["meta"][f"bridson_{name}"]["args"]
        rc, out = cli("bridson_points", dims, radius, seed, f"{name}.npy", cwd=tmp_path)
        assert rc == 0 and out.strip() == golden["meta"][f"bridson_{name}"]["stdout"]
        got = np.load(tmp_path / f"{name}.npy")
        assert np.array_equal(got.view(np.uint32), golden["next"][f"bridson_{name}"].view(np.uint32))


def _error_case_inputs(d):
    np.save(d / "pts3.npy", np.zeros((4, 3), np.float32))
    np.save(d / "pts64.npy", np.zeros((4, 2), np.float64))
    np.save(d / "pts1d.npy", np.zeros(8, np.float32))
    np.save(d / "vel2d.npy", np.zeros((8, 8), np.float32))
    np.save(d / "vel64.npy", np.zeros((8, 8, 2), np.float64))
    np.save(d / "ok_pts.npy", np.ones((4, 2), np.float32))
    np.save(d / "ok_vel.npy", np.zeros((8, 8, 2), np.float32))
    np.save(d / "img3d.npy", np.zeros((4, 4, 2), np.float32))
    np.save(d / "img64.npy", np.zeros((4, 4), np.float64))


def test_argument_errors_of_the_headline_commands_match_the_reference(cli, golden, tmp_path):
    """advect_points.cc:65-68,84-104, splat_points.cc:45-100, generate_simplex.cc:44-47, curl_2d.cc:30-45: the
    reference's own stdout and exit code for each bad invocation (tests/golden/make_golden.py ran the reference
    CLI on the same files).  All of these are decided on the host, before a device is needed."""
    _error_case_inputs(tmp_path)
    cases = golden["meta"]["cli_errors"]
    assert len(cases) >= 17
    for c in cases:
        rc, out = cli(*c["args"], cwd=tmp_path)
        assert rc == c["rc"] == 1, c
        assert out.strip().replace("\r", "\n") == c["stdout"].replace("\r", "\n"), (c["args"], out)
        assert not (tmp_path / "x.npy").exists()


def test_device_commands_fail_loudly_without_a_gpu(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, out = cli("pendulum_phase", "64x48", "0.3", "0.7", "1.3", "f.npy", cwd=tmp_path)
    assert rc == 1 and "clumpy_cuda" in out and not (tmp_path / "f.npy").exists()


@pytest.mark.gpu
def test_pendulum_phase_file_is_byte_identical_to_the_reference(cli, golden, tmp_path):
    """README example6's field: pendulum_phase 4000x2000 0.9 2 5 (md5 from SURVEY.md 8c)."""
    rc, out = cli("pendulum_phase", "4000x2000", "0.9", "2", "5", "field.npy", cwd=tmp_path)
    assert rc == 0, out
    assert md5(tmp_path / "field.npy") == golden["meta"]["c2_md5"]["field.npy"] == "c6107b1fa21605d0a8c920f02b6e6710"
    for name in ("p0", "p1", "p2"):
        dims, friction, sx, sy = golden["meta"][f"pendulum_{name}"]["args"]
        rc, out = cli("pendulum_phase", dims, friction, sx, sy, f"{name}.npy", cwd=tmp_path)
        assert rc == 0, out
        got = np.load(tmp_path / f"{name}.npy")
        assert np.array_equal(got.view(np.uint32), golden["next"][f"pendulum_{name}"].view(np.uint32))


@pytest.mark.gpu
def test_cull_points_cli_matches_the_reference(cli, golden, tmp_path):
    nxt = golden["next"]
    np.save(tmp_path / "pts.npy", nxt["cull_pts"])
    np.save(tmp_path / "mask.npy", nxt["cull_mask"])
    rc, out = cli("cull_points", "pts.npy", "mask.npy", "culled.npy", cwd=tmp_path)
    assert rc == 0 and out.strip() == golden["meta"]["cull"]["stdout"]
    got = np.load(tmp_path / "culled.npy")
    assert got.shape == nxt["cull_out"].shape and np.array_equal(got.view(np.uint32), nxt["cull_out"].view(np.uint32))
    # wrong dtype / shape: the reference's messages
    np.save(tmp_path / "bad.npy", nxt["cull_pts"].astype(np.float64))
    rc, out = cli("cull_points", "bad.npy", "mask.npy", "x.npy", cwd=tmp_path)
    assert rc == 1 and out.strip() == "Input data has wrong data type."
    np.save(tmp_path / "bad3.npy", np.zeros((4, 3), np.float32))
    rc, out = cli("cull_points", "bad3.npy", "mask.npy", "x.npy", cwd=tmp_path)
    assert rc == 1 and out.strip() == "Input points must be 2D."


# ---- the four headline commands through the CLI on a GPU: file bytes, names and stdout ---------------------------

@pytest.mark.gpu
def test_c1_pipeline_files_are_byte_identical_to_the_reference(cli, golden, oracle_mod, tmp_path):
    """README example5 through `clumpy curl_2d` and `clumpy advect_points` (k = 1: frames must be bit-identical):
    every output FILE -- .npy header bytes included -- has the md5 of the reference CLI's file, the 240 frames
    are named {:03}{suffix} (advect_points.cc:177-178) and stdout is the reference's byte for byte.
    The potential fed to curl_2d is the reference's (the CPU restatement, bit-identical to it): the CUDA noise
    field agrees with it to 1e-5, not to the bit, so its own file is checked separately below."""
    m = golden["meta"]
    pot, _, _ = oracle_mod.generate_simplex(1000, 500, 1.0, 8.0, 0)
    np.save(tmp_path / "potential.npy", pot)
    np.save(tmp_path / "pts.npy", golden["c1_pts"])
    rc, out = cli("curl_2d", "potential.npy", "velocity.npy", cwd=tmp_path)
    assert rc == 0 and out.strip().splitlines() == m["c1_curl_stdout"]
    assert md5(tmp_path / "velocity.npy") == m["c1_md5"]["velocity.npy"] == "affc7ee1fad17db42f058c56bb54b4df"
    rc, out = cli("advect_points", "pts.npy", "velocity.npy", "30", "1", "0.95", "240", "anim.npy", cwd=tmp_path)
    assert rc == 0
    assert hashlib.md5(out.encode()).hexdigest() == m["c1_advect_stdout_md5"]
    assert out.endswith("\nGenerated 000anim.npy through 239anim.npy.\n")
    names = sorted(p.name for p in tmp_path.glob("*anim.npy"))
    assert names == [f"{i:03}anim.npy" for i in range(240)]
    assert [md5(tmp_path / n) for n in names] == m["c1_frame_md5"]
    assert m["c1_frame_md5"][0] == "03c381895efdd94ac301ef66906fe9f0" and m["c1_frame_md5"][239] == "91c4ede7beb0c9f41499cf3831c341b0"


@pytest.mark.gpu
def test_generate_simplex_cli(cli, golden, tmp_path):
    """generate_simplex.cc:43-87 through the CLI: same .npy header bytes and shape as the reference's file, data
    within 1e-5, printed range equal to the reference's line up to the last printed digit."""
    m = golden["meta"]
    rc, out = cli("generate_simplex", "1000x500", "1.0", "8.0", "0", "potential.npy", cwd=tmp_path)
    assert rc == 0 and out.startswith("Noise range is ")
    got_lo, got_hi = (float(v) for v in out.replace("Noise range is ", "").split(" to "))
    ref_lo, ref_hi = (float(v) for v in m["c1_simplex_stdout"].replace("Noise range is ", "").split(" to "))
    assert abs(got_lo - ref_lo) <= 2e-6 and abs(got_hi - ref_hi) <= 2e-6
    raw = open(tmp_path / "potential.npy", "rb").read()
    assert len(raw) == 80 + 1000 * 500 * 4  # the header of a (500, 1000) array is 80 bytes
    assert raw[:10] == b"\x93NUMPY\x01\x00\x46\x00" and raw[79:80] == b"\n"
    assert raw[10:80].rstrip(b" \n") == b"{'descr': '<f4', 'fortran_order': False, 'shape': (500, 1000), }"
    for name in ("a", "b", "c"):
        dims, amp, freq, seed = m[f"simplex_{name}"]["args"]
        rc, out = cli("generate_simplex", dims, amp, freq, seed, f"pot_{name}.npy", cwd=tmp_path)
        assert rc == 0
        got = np.load(tmp_path / f"pot_{name}.npy")
        assert np.abs(got - golden["fields"][f"pot_{name}"]).max() <= 1e-5 * max(1.0, float(amp))
        np.save(tmp_path / f"ref_pot_{name}.npy", golden["fields"][f"pot_{name}"])
        rc, out = cli("curl_2d", f"ref_pot_{name}.npy", f"vel_{name}.npy", cwd=tmp_path)
        assert rc == 0 and out.strip().splitlines() == m[f"curl_{name}"]["stdout"]
        assert np.array_equal(np.load(tmp_path / f"vel_{name}.npy").view(np.uint32), golden["fields"][f"vel_{name}"].view(np.uint32))


@pytest.mark.gpu
def test_splat_points_cli(cli, golden, tmp_path):
    """splat_points.cc:44-104 through the CLI: stdout, u8 .npy header bytes, the alpha = 1 / k = 1 fast path."""
    m, spl = golden["meta"], golden["splat"]
    np.save(tmp_path / "pts.npy", spl["pts"])
    for name in ("k1_a1", "k3_a1", "k5_a05", "k1_a03", "k7_a1"):
        c = m[f"splat_{name}"]
        rc, out = cli("splat_points", "pts.npy", "96x64", "u8disk", c["k"], c["alpha"], f"{name}.npy", cwd=tmp_path)
        assert rc == 0 and out.strip() == c["stdout"] == f"Drawing {len(spl['pts'])} points."
        raw = open(tmp_path / f"{name}.npy", "rb").read()
        assert raw[:8] == b"\x93NUMPY\x01\x00" and b"'descr': '<u1'" in raw[:80] and b"'shape': (64, 96), }" in raw[:80]
        assert (len(raw) - 64 * 96) % 16 == 0
        got = np.load(tmp_path / f"{name}.npy")
        d = np.abs(got.astype(int) - spl[name].astype(int))
        assert (d <= 1).mean() >= 0.999 and d.max() <= 1, name
        if c["k"] == 1:
            assert np.array_equal(got, spl[name]), name
    # fp32sphere takes the fast path when alpha = 1 and k = 1 (splat_points.cc:93-94), and is refused otherwise
    rc, out = cli("splat_points", "pts.npy", "96x64", "fp32sphere", "1", "1.0", "sphere.npy", cwd=tmp_path)
    assert rc == 0 and np.array_equal(np.load(tmp_path / "sphere.npy"), spl["k1_a1"])


@pytest.mark.gpu
def test_advect_points_cli_frames_of_the_small_fixtures(cli, golden, tmp_path):
    """advect_points through the CLI on the small golden runs (k = 1, 3, x-wrap via a negative decay, decay 0):
    frame files compared with the reference CLI's frames."""
    adv, m = golden["advect"], golden["meta"]
    np.save(tmp_path / "pts.npy", adv["pts"])
    for vname in ("vel_small", "vel_drift"):
        np.save(tmp_path / f"{vname}.npy", adv[vname])
    for name in ("k1", "k3", "k3_wrap", "k1_decay0", "k5"):
        c = m[f"advect_{name}"]
        rc, out = cli("advect_points", "pts.npy", f"{c['vel']}.npy", c["step"], c["k"], c["decay"], c["nframes"], f"_{name}.npy", cwd=tmp_path)
        assert rc == 0, out
        ref = adv[f"frames_{name}"]
        for f in range(c["nframes"]):
            raw = open(tmp_path / f"{f:03}_{name}.npy", "rb").read()
            assert b"'descr': '<u1'" in raw[:80] and b"'shape': (64, 96), }" in raw[:80]
            d = np.abs(np.load(tmp_path / f"{f:03}_{name}.npy").astype(int) - ref[f].astype(int))
            assert d.max() <= 1 and (d == 0).mean() >= 0.998, (name, f)
            if c["k"] == 1:
                assert d.max() == 0, (name, f)
