"""bench.py's JSON contract, checked on the CPU through its reference arm (the only arm that runs without a GPU):
one line, the contract keys, the configuration both arms share (incl. the CPU sample), `impl: reference`,
zero copy bytes.  The GPU arm prints the same keys plus roofline / cpu_baseline / clocks / gpu_launches (checked by
the driver on the B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bits", "512", "--n", "300",
           "--cpu-sample", "24", "--steps", "1", "--warmup", "0", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    line = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "ciphertexts/s" and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["vs_baseline"] is None and line["value"] > 0
    cfg = line["config"]
    assert "N=300 ciphertexts in total" in cfg["workload"] and cfg["n_total"] == 300 and cfg["cpu_sample"] == 24
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["sample"].startswith("24 of 300 ciphertexts")
    assert line["e2e"] == {"value": line["value"], "unit": "ciphertexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_verify_mix_line():
    line = _run("--workload", "verify-mix")
    assert line["impl"] == "reference" and "verification of a 3-party mix" in line["config"]["workload"]
    assert line["config"]["cpu_sample"] == 24 and line["value"] > 0


def test_reference_arm_committed_shuffle_line():
    line = _run("--workload", "committed-shuffle", "--width", "3")
    assert line["impl"] == "reference" and "commitment-consistent proof of a shuffle" in line["config"]["workload"]
    assert line["config"]["cpu_sample"] == 24 and line["config"]["width"] == 3 and line["value"] > 0
    assert line["cpu_baseline"]["sample"].startswith("24 of 300 ciphertexts of width 3")


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
