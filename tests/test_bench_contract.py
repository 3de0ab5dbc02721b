"""bench.py contract checks that need no GPU: the reference arm runs the unmodified reference's
CPU path and prints one JSON line with the required keys; the default arm refuses to run without
a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--size", str(4 << 20)],      # the default (the genuine 64 MiB block) takes ~90 s
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-800:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MB/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.1 < d["value"] < 1000
    assert d["cpu_baseline"]["cores"] == 1 and d["steps_measured"] == 1 and d["all_cores_slices"]["cores"] >= 1
    assert d["config"]["block_bytes"] == 4 << 20
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    """the evidence under profiles/ is the output of bench.py: every key the contract names is there"""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "roofline_decompress", "cpu_baseline", "clocks",
              "calgary_batch", "single_block_1g"):
        assert k in d, k
    assert d["config"]["workload"] == "text64m" and d["file_sha256_matches_reference_golden"] is True
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["text"][str(1 << 30)]
    secs = []
    for n in (1, 2, 4, 8):
        b = json.load(open(os.path.join(ROOT, "profiles", "scaling", "r2_bench_n%d.json" % n)))
        assert b["n_gpus"] == n and b["single_block_1g"]["n_gpus"] == n
        assert b["single_block_1g"]["sha256"] == golden["sha256"] and b["single_block_1g"]["matches_golden"] is True
        secs.append(b["single_block_1g"]["seconds"])
    assert secs == sorted(secs, reverse=True), secs        # the one 1 GiB block gets faster with every doubling of GPUs
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_reference.json")))
    assert ref["impl"] == "reference" and ref["config"]["block_bytes"] == d["config"]["block_bytes"] and ref["cpu_baseline"]["cores"] == 1
