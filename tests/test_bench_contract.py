"""-m "not gpu": the reference arm of bench.py (`--impl reference`: the oracle port of the
generator on the host cores) runs without a GPU and prints ONE JSON line with the contract's
keys; the B200 arm refuses to run without its CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "generated audio samples/sec"
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "BASELINE config 3" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_b200_arm_fails_loudly_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1",
                        "--warmup", "0", "--no-cpu-baseline"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())


def test_only_the_cpu_baseline_leg_of_bench_touches_the_oracle():
    """bench.py may execute oracle/ only as the reported CPU baseline (`cpu_port`, shared by the
    cpu_baseline key and the `--impl reference` arm): no other function imports it."""
    import ast

    def oracle_imports(node):
        for n in ast.walk(node):
            if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle":
                yield n
            if isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names):
                yield n

    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    owners = set()
    for top in tree.body:
        if list(oracle_imports(top)):
            owners.add(getattr(top, "name", "<module level>"))
    assert owners == {"cpu_port"}, owners
