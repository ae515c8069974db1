"""The reference arm of bench.py (`--impl reference`) needs no GPU: check its one-line JSON contract
here, on BASELINE configs[0] (the reference's own CPU-runnable case)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1_uniform4096_n64_fp32", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_json_contract():
    lines = _run()
    assert len(lines) == 1                                   # exactly ONE JSON line
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GFLOP/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("SpMM GFLOP/s") and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["config"]["workload"] == "cfg1_uniform4096_n64_fp32" and j["config"]["n"] == 64
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "rows" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None and j["data"] == "synthetic"


def test_reference_arm_only_rank0_works_under_torchrun_env():
    # ranks != 0 exit 0 without output (the driver launches the arm with torchrun for N > 1)
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
