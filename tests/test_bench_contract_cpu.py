"""The `bench.py --impl reference` arm (the reference's algorithm on the host cores: oracle port, no GPU) prints exactly
ONE JSON line on stdout with the keys the driver's contract names; native-library chatter goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, COMMU_CPU_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0
    assert j["metric"].startswith("train tokens/sec @12L d512 seq2048") and j["unit"] == "tokens/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
