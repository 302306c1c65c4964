#!/bin/bash
# ncu --set full on the sparse-conv micro-benchmark (few launches, cheap)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"spconv_tc_kernel|spconv_wgrad_tc_kernel" -s ${SKIP:-0} -c ${COUNT:-8} -f -o gpurun_out/prof_conv \
    python scripts/bench_conv.py fp32x3 > gpurun_out/ncu_conv.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_conv.ncu-rep; tail -3 gpurun_out/ncu_conv.log
