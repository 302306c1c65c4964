#!/bin/bash
# Round-trip used during development: GPU parity tests, then a memcheck pass over a small subset.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -40
if [ "$1" == "sanitize" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu -x \
    "tests/test_gpu_voxelize.py::test_voxelize_ragged_and_empty_scenes" \
    "tests/test_gpu_spconv.py::test_sparse_rulebook_exact" \
    "tests/test_gpu_spconv.py::test_dense_roundtrip_and_grad" \
    "tests/test_gpu_box_attn.py::test_box_attn_matches_reference_golden" \
    -k "not full_size" > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer exit: $?"; tail -15 gpurun_out/sanitizer.log
fi
