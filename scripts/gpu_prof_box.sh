#!/bin/bash
# ncu --set full capture of the box-attention kernels (unfused and fused) at the encoder geometry; CSV pages come back.
mkdir -p gpurun_out
TAG=${TAG:-box}
ITERS=1 WARM=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"box_attn_(fwd|bwd)_tile_kernel" -f -o /tmp/prof_$TAG \
    python scripts/bench_box_attn.py > gpurun_out/ncu_$TAG.log 2>&1
echo "capture rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
python scripts/ncu_summarize.py raw gpurun_out/prof_${TAG}_raw.csv gpurun_out/prof_${TAG}_summary.txt
rm -f gpurun_out/prof_${TAG}_raw.csv
cat gpurun_out/prof_${TAG}_summary.txt | head -150
