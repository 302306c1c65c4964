#!/bin/bash
# Round-2 evidence run on ONE GPU: ncu launch list of one (eager) step with DRAM bytes, ncu --set full of the sparse-conv
# kernels on the micro-benchmark geometry, compute-sanitizer memcheck + racecheck over the tcgen05 / box-attention tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-step --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python scripts/ncu_summarize.py launches gpurun_out/launches.csv gpurun_out/r2_launches_step.txt gpurun_out/r2_traffic_step.json; head -30 gpurun_out/r2_launches_step.txt
MODE=bf16x3 LEVELS=0,1,2,3 KINDS=fwd,wgrad TAG=r2conv bash scripts/gpu_prof_conv.sh > gpurun_out/prof_conv.log 2>&1; tail -5 gpurun_out/prof_conv.log
rm -f gpurun_out/prof_r2conv_raw.csv gpurun_out/prof_r2conv_source.csv
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_spconv_tc.py tests/test_gpu_box_attn.py -m gpu -q -x --timeout 500 \
      -k "forward_vs_oracle and (16-16-3000 or 64-64-127 or 128-128-2000) or wgrad or golden or module_forward_backward" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|hazard" gpurun_out/sanitizer_$tool.log | tail -5
done
du -sh gpurun_out
