#!/bin/bash
# ncu --set full capture of the sparse-conv kernels on the micro-benchmark geometry; CSV pages come back, not the report.
# usage: MODE=bf16x3 LEVELS=0,1 KINDS=fwd TAG=r2a bash scripts/gpu_prof_conv.sh
mkdir -p gpurun_out
TAG=${TAG:-conv}
ITERS=1 WARM=0 LEVELS=${LEVELS:-0,1} KINDS=${KINDS:-fwd} timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"spconv_tc_kernel|spconv_wgrad_tc_kernel" -f -o /tmp/prof_$TAG python scripts/bench_conv.py ${MODE:-bf16x3} > gpurun_out/ncu_$TAG.log 2>&1
echo "capture rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
python scripts/ncu_summarize.py raw gpurun_out/prof_${TAG}_raw.csv gpurun_out/prof_${TAG}_summary.txt
python scripts/ncu_source_top.py gpurun_out/prof_${TAG}_source.csv 40 > gpurun_out/prof_${TAG}_hotspots.txt 2>&1
head -60 gpurun_out/prof_${TAG}_summary.txt; head -50 gpurun_out/prof_${TAG}_hotspots.txt
