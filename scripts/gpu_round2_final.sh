#!/bin/bash
# Final evidence run of round 2 on ONE GPU: full GPU suite, bench lines of every BASELINE config, ncu launch list of one
# eager step (DRAM bytes per launch), micro-benchmarks, compute-sanitizer memcheck over the kernels touched last.
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_voxel_detr.err | tail -1 > $O/bench_voxel_detr.json; echo "bench rc=$?"
for w in conquer centerpoint_waymo centerpoint_nusc config1; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 2>$O/bench_$w.err | tail -1 > $O/bench_$w.json; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for w in ("voxel_detr", "conquer", "centerpoint_waymo", "centerpoint_nusc", "config1"):
    try:
        d = json.load(open("gpurun_out/final/bench_%s.json" % w))
        print(w, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(w, "FAILED", e)
PY
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/launches.csv python bench.py --profile-step --warmup 3 > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l $O/launches.csv
python scripts/ncu_summarize.py launches $O/launches.csv $O/r2_launches_step_eager.txt $O/r2_traffic_step.json; head -24 $O/r2_launches_step_eager.txt
rm -f $O/launches.csv
timeout 200 python scripts/bench_conv.py bf16x3 > $O/conv_micro.txt 2>&1; tail -14 $O/conv_micro.txt
timeout 100 python scripts/bench_dense.py > $O/dense_micro.txt 2>&1; tail -3 $O/dense_micro.txt
timeout 100 python scripts/bench_box_attn.py > $O/box_micro.txt 2>&1; tail -5 $O/box_micro.txt
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_spconv_tc.py tests/test_gpu_box_attn.py tests/test_gpu_batchnorm.py -m gpu -q -x --timeout 600 \
    -k "forward_vs_oracle and (16-16-3000 or 64-64-127 or 128-128-2000) or refresh_packs or fused_where or (module_forward_backward and (5-16 or 16-16)) or golden or bn" > $O/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -4
du -sh $O
