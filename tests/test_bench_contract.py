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
    # sparse-conv launches and dense-linear launches of one __global__ function are separate roofline groups
    assert bench.kernel_of("spconv_tc_c64") == "spconv_tc_kernel[sparse conv]"
    assert bench.kernel_of("dense_tc_gemm") == "spconv_tc_kernel[dense linear]"
    assert bench.kernel_of("spconv_tc_wgrad_c128") == "spconv_wgrad_tc_kernel[sparse conv]"
    assert bench.kernel_of("dense_tc_wgrad") == "spconv_wgrad_tc_kernel[dense linear]"
    assert bench.kernel_of("box_attn_bwd") == "box_attn_bwd_tile_kernel"
    assert bench.kernel_of("spconv_gemm_c16") == "spconv_fwd_kernel"
    assert bench.kernel_of("lsa") == "lsa"


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    traffic, src = bench.ncu_traffic("spconv_tc_kernel[sparse conv]")
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
    assert a.workload == "voxel_detr" and a.scenes == 2 and a.points == 150000 and a.workload_name == bench.WORKLOAD
    p = bench.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["bf16_tflops"] > 100 and p["source"] in ("measured", "fallback")


def test_clock_sampler_without_a_gpu_reports_why():
    s = bench.ClockSampler(0)
    s.start()
    s.stop()
    r = s.result()
    assert "reasons" in r and "sm_mhz" in r


def test_rooflines_keep_sparse_and_dense_launches_apart():
    peaks = {"hbm_gbs": 6555.5, "bf16_tflops": 1389.0, "source": "measured"}
    summary = {"spconv_tc_c16": {"launches": 2, "ms": 0.1, "bytes": 70e6, "flops": 4e9},
               "dense_tc_gemm": {"launches": 4, "ms": 0.6, "bytes": 1.2e9, "flops": 3e11},
               "voxelize": {"launches": 1, "ms": 0.06, "bytes": 14e6, "flops": 0}}
    rl = bench.rooflines(summary, 1, 10.0, peaks)
    sp, de = rl["spconv_tc_kernel[sparse conv]"], rl["spconv_tc_kernel[dense linear]"]
    assert sp["bound"] == "hbm" and abs(sp["achieved"] - 700.0) < 1 and abs(sp["frac"] - 700.0 / 6555.5) < 1e-3
    assert de["bound"] == "tensor" and abs(de["achieved"] - 500.0) < 1
    assert rl["voxelize"]["bound"] == "hbm"
