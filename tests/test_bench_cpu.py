"""The CPU arm of bench.py (`--impl reference`): the driver runs it on the GPU box beside the GPU arm and divides the two
lines, so its JSON contract is pinned here -- no GPU needed, the arm times the compiled reference (oracle/_ref) or the
oracle port on the host cores."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout  # the contract: ONE JSON line on stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("extra", [(), ("--workload", "c3")])
def test_reference_arm_prints_one_contract_line(extra):
    line = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", *extra)
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    for key in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in line, key
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert "workload" in line["config"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["sample"] and cpu["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
