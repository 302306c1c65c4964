#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -E "^E  |passed|failed|Error|FAILED" | head -30 | tee gpurun_out/pytest_gpu.log
python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_line.json
