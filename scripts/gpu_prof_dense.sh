#!/bin/bash
# ncu --set full capture (warm caches) of the dense token-wise linear on the tcgen05 kernel; CSV pages come back.
mkdir -p gpurun_out
TAG=${TAG:-dense}
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"spconv_tc_kernel" -s 4 -c 2 -f -o /tmp/prof_$TAG \
    python scripts/bench_dense.py > gpurun_out/ncu_$TAG.log 2>&1
echo "capture rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
python scripts/ncu_summarize.py raw gpurun_out/prof_${TAG}_raw.csv gpurun_out/prof_${TAG}_summary.txt
python scripts/ncu_source_top.py gpurun_out/prof_${TAG}_source.csv 60 > gpurun_out/prof_${TAG}_hotspots.txt 2>&1
head -40 gpurun_out/prof_${TAG}_summary.txt; head -75 gpurun_out/prof_${TAG}_hotspots.txt
