#!/bin/bash
# GPU round-trip: model parity tests, smoke, a short bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_model.log; tail -30 gpurun_out/pytest_model.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_line.json
