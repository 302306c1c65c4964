"""CPU: the pieces of bench.py that do not need a GPU — launch-family -> kernel mapping, the roofline traffic
lookup in the committed ncu summary, clock-sampler fallback, argument defaults."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_family_to_kernel_mapping():
    assert bench.kernel_of("spconv_tc_c64") == "spconv_tc_kernel"
    assert bench.kernel_of("dense_tc_gemm") == "spconv_tc_kernel"
    assert bench.kernel_of("spconv_tc_wgrad_c128") == "spconv_wgrad_tc_kernel"
    assert bench.kernel_of("dense_tc_wgrad") == "spconv_wgrad_tc_kernel"
    assert bench.kernel_of("box_attn_bwd") == "box_attn_bwd_tile_kernel"
    assert bench.kernel_of("spconv_gemm_c16") == "spconv_fwd_kernel"
    assert bench.kernel_of("lsa") == "lsa"


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    traffic, src = bench.ncu_traffic("spconv_tc_kernel")
    assert src is not None and src.startswith("profiles/") and os.path.exists(os.path.join(ROOT, src))
    with open(os.path.join(ROOT, src)) as f:
        js = json.load(f)
    entry = [d for k, d in js["kernels"].items() if k.split("::")[-1] == "spconv_tc_kernel"][0]
    assert traffic == entry["dram_bytes_per_launch"] > 0
    assert bench.ncu_traffic("no_such_kernel") == (None, None)


def test_defaults_and_peaks():
    argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse_args()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.steps >= 1 and a.warmup >= 3 and a.impl == "efgb200"
    p = bench.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["bf16_tflops"] > 100 and p["source"] in ("measured", "fallback")


def test_clock_sampler_without_a_gpu_reports_why():
    s = bench.ClockSampler(0)
    s.start()
    s.stop()
    r = s.result()
    assert "reasons" in r and "sm_mhz" in r
