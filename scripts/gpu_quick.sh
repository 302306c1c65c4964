#!/bin/bash
# Short GPU round-trip: parity tests, bench (default + optional A/B environments / library variants), micro-benchmarks.
mkdir -p gpurun_out
timeout ${TEST_TIMEOUT:-300} python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench_line.json
echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_line.json; tail -3 gpurun_out/bench.err
timeout 120 python scripts/bench_conv.py fp32x3 > gpurun_out/conv_micro.txt 2>&1; tail -9 gpurun_out/conv_micro.txt
timeout 120 python scripts/bench_dense.py > gpurun_out/dense_micro.txt 2>&1; tail -12 gpurun_out/dense_micro.txt
i=0
for e in ${ALT_ENVS}; do
  i=$((i+1))
  env $e timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_alt$i.err | tail -1 > gpurun_out/bench_line_alt$i.json
  echo "bench [$e] rc=$?"; cut -c1-260 gpurun_out/bench_line_alt$i.json; tail -3 gpurun_out/bench_alt$i.err
  env $e timeout 120 python scripts/bench_conv.py fp32x3 > gpurun_out/conv_micro_alt$i.txt 2>&1; echo "[$e]"; tail -9 gpurun_out/conv_micro_alt$i.txt
  env $e timeout 120 python scripts/bench_dense.py > gpurun_out/dense_micro_alt$i.txt 2>&1; tail -12 gpurun_out/dense_micro_alt$i.txt
done
if [ "${BOX:-0}" == "1" ]; then timeout 60 python scripts/bench_box_attn.py > gpurun_out/box_micro.txt 2>&1; cat gpurun_out/box_micro.txt; fi
timeout 120 python scripts/prof_step.py > gpurun_out/prof_step.txt 2>&1; head -2 gpurun_out/prof_step.txt
if [ "${PHASES:-0}" == "1" ]; then timeout 150 python scripts/prof_phases.py > gpurun_out/prof_phases.txt 2>&1; head -20 gpurun_out/prof_phases.txt; fi
