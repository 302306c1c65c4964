#!/bin/bash
# Short GPU round-trip: parity tests, bench (default + optional A/B variant), micro-benchmarks.  No ncu.
mkdir -p gpurun_out
timeout ${TEST_TIMEOUT:-300} python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench_line.json
echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_line.json; tail -3 gpurun_out/bench.err
for v in ${VARIANTS}; do
  EFGB_LIB_VARIANT=$v timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$v.err | tail -1 > gpurun_out/bench_line_$v.json
  echo "bench variant $v rc=$?"; cut -c1-260 gpurun_out/bench_line_$v.json; tail -3 gpurun_out/bench_$v.err
  EFGB_LIB_VARIANT=$v timeout 120 python scripts/bench_conv.py fp32x3 > gpurun_out/conv_micro_$v.txt 2>&1; tail -9 gpurun_out/conv_micro_$v.txt
done
timeout 60 python scripts/bench_box_attn.py > gpurun_out/box_micro.txt 2>&1; cat gpurun_out/box_micro.txt
timeout 120 python scripts/bench_conv.py fp32x3 > gpurun_out/conv_micro.txt 2>&1; tail -9 gpurun_out/conv_micro.txt
timeout 120 python scripts/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -2 gpurun_out/prof_step.txt
