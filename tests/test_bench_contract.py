"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm's JSON line (the unmodified
reference through the same streaming host), identical `config` objects for both arms, and no CPU fallback in the CUDA arm."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "kvz_stream_bench_ref")


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_line_has_the_contract_keys():
    if not os.path.exists(REF_BENCH):
        pytest.skip("oracle/_ref/kvz_stream_bench_ref missing (make -C integration needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "64x64", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # ONE JSON line
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "encoded frames/sec at fixed QP (bit-identical bitstream)" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    b = _bench_module()
    assert d["config"] == b.config_of(b.WORKLOADS["64x64"])  # what the CUDA arm reports for the same workload


def test_default_workload_is_the_headline_config():
    b = _bench_module()
    wl = b.WORKLOADS["2160p"]
    assert (wl["w"], wl["h"], wl["preset"], wl["qp"]) == (3840, 2160, "veryslow", 22)        # BASELINE config 3
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--workload", default="2160p"' in src


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for boxes without a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "64x64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0                                  # no CPU fallback
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
