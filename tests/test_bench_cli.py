"""CPU: bench.py's command line -- the reference arm (the oracle port on the host cores) prints the contract's JSON line,
and the product arm refuses to run without a GPU instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "A", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "POI/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("A:")
    # the CPU arm names its CPU, its core and thread counts, and reports best-of-N next to the mean
    cb = d["cpu_baseline"]
    assert cb["cpu_model"] and cb["logical_cpus"] >= cb["cores"] and cb["value"] >= cb["value_mean"] > 0
    assert "OMP_PROC_BIND=close" in cb["sample"] and "best of" in cb["sample"]
    assert set(d["config"]) == {"workload", "pois_per_gpu", "conv", "stop", "parallelism", "l2"}  # the GPU arm's keys


def test_product_arm_needs_a_gpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "A", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stdout + out.stderr
