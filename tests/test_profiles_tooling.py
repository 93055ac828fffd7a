"""CPU: the committed ncu evidence under profiles/ is parseable by the repo's own tools and consistent with itself."""
import gzip
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


import pytest


@pytest.mark.parametrize("raw,summary", [("r01_launches_s9.csv.gz", "r01_launches_step_summary_s9.json"),
                                         ("r02_launches.csv.gz", "r02_launches_step_summary.json")])
def test_launch_list_summarises_to_the_committed_step_summary(tmp_path, raw, summary):
    csv_path = tmp_path / "launches.csv"
    with gzip.open(os.path.join(PROF, raw), "rb") as f:
        csv_path.write_bytes(f.read())
    out = tmp_path / "summary.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), str(csv_path), str(out)], check=True,
                   capture_output=True)
    got = json.load(open(out))
    want = json.load(open(os.path.join(PROF, summary)))
    assert "step" in got["window"] and got["launches_in_step"] == want["launches_in_step"]
    assert abs(sum(k["share"] for k in got["kernels"]) - 1.0) < 1e-2
    names = {k["kernel"]: k for k in got["kernels"]}
    wnames = {k["kernel"]: k for k in want["kernels"]}
    assert set(names) == set(wnames)
    for n, k in names.items():
        assert k["launches"] == wnames[n]["launches"] and abs(k["total_ms"] - wnames[n]["total_ms"]) < 1e-3
    # one OPT-1.3B step: 24 layers x (2 norm+quantize, 6 Linears, 1 attention) + lm_head + loss
    ours = {n: k["launches"] for n, k in names.items() if n.startswith("bq::")}
    assert sum(v for n, v in ours.items() if "gemm_bf16_tn_kernel<256, 0, 2>" not in n and "gemm_bf16_tn_kernel" in n) == 144
    assert sum(v for n, v in ours.items() if "attention_causal" in n) == 24          # round 2: attention_causal_dual_kernel
    assert sum(v for n, v in ours.items() if "norm_quant" in n) == 48


def test_bench_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench

    traffic, src = bench.gemm_traffic_from_profile()
    assert src == os.path.join("profiles", "r02_ncu_layer.json")
    assert 2.0e8 < traffic < 4.0e8                      # ~290 MB per launch against 285 MB algorithmic
    head = [json.loads(x) for x in open(os.path.join(PROF, "r02_bench_n1_s3.json")).read().strip().splitlines() if x.startswith("{")][-1]
    r = head["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["traffic"] - traffic) / traffic < 0.05
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "gpu_port_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in head, k
    assert head["gpu_launches"] > 0 and head["e2e"]["h2d_bytes_per_step"] == 8 * 2048 * 8 and head["e2e"]["d2h_bytes_per_step"] == 4
