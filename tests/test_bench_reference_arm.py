"""CPU: bench.py's reference arm (oracle/_ref -- the reference's own decoder sources -- on host
cores, or the oracle port where that library is not built) prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gbit/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1
    from oracle import ref as R
    assert line["cpu_baseline"]["kind"] == ("reference" if R.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    if R.available():                       # the same-work figure of the port is reported beside it
        assert 0 < line["cpu_baseline"]["port_same_work"]["value"] <= line["value"] * 1.5
    assert line["e2e"] == {"value": line["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
