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
        p = subprocess.run([CLI] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)
        return p.returncode, p.stdout
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
