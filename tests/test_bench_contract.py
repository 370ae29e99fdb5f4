"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints one JSON line with the
keys the driver reads, and the product arm refuses to run on a CPU (no fallback path exists)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # the same metric and workload as the product arm reports (BASELINE.json's north star)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "metric" in base and d["metric"]


def test_product_arm_refuses_to_run_without_a_gpu():
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stdout + res.stderr)
    assert not [ln for ln in res.stdout.splitlines() if ln.strip().startswith("{")], "no result line may be printed"
