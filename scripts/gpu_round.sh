#!/bin/bash
# One GPU round-trip: parity tests, bench line(s), ncu launch list of ONE step, ncu --set full of the hot kernels
# on the micro-benchmarks (converted to CSV on the box: a .ncu-rep with sources is too big to travel back).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${NOTEST:-0}" != "1" ]; then
timeout ${TEST_TIMEOUT:-300} python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
fi
timeout 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench_line.json
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_line.json; tail -3 gpurun_out/bench.err
if [ -n "${ALT_ENV}" ]; then
env ${ALT_ENV} timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_alt.err | tail -1 > gpurun_out/bench_line_alt.json
echo "alt bench (${ALT_ENV}) rc=$?"; cut -c1-300 gpurun_out/bench_line_alt.json
fi
timeout 60 python scripts/bench_box_attn.py > gpurun_out/box_micro.txt 2>&1; cat gpurun_out/box_micro.txt
timeout 120 python scripts/bench_dense.py > gpurun_out/dense_micro.txt 2>&1; tail -4 gpurun_out/dense_micro.txt
timeout 120 python scripts/bench_conv.py fp32x3 > gpurun_out/conv_micro.txt 2>&1; tail -9 gpurun_out/conv_micro.txt
timeout 150 python scripts/prof_phases.py > gpurun_out/prof_phases.txt 2>&1; head -16 gpurun_out/prof_phases.txt
if [ "${NONCU:-0}" != "1" ]; then
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile-step --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
ITERS=1 WARM=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"box_attn" -f -o /tmp/prof_box \
    python scripts/bench_box_attn.py > gpurun_out/ncu_box.log 2>&1
echo "box capture rc=$?"
ITERS=1 WARM=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"spconv_tc_kernel|spconv_wgrad_tc_kernel" -f -o /tmp/prof_conv \
    python scripts/bench_conv.py fp32x3 > gpurun_out/ncu_conv.log 2>&1
echo "conv capture rc=$?"
for n in box conv; do
  if [ -f /tmp/prof_$n.ncu-rep ]; then
    ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_$n.ncu-rep --page details --csv > gpurun_out/prof_${n}_details.csv 2>/dev/null
    ncu -i /tmp/prof_$n.ncu-rep --page source --csv > gpurun_out/prof_${n}_source.csv 2>/dev/null
    sz=$(stat -c %s /tmp/prof_$n.ncu-rep); echo "prof_$n.ncu-rep $sz bytes"
    if [ $sz -lt 20000000 ]; then cp /tmp/prof_$n.ncu-rep gpurun_out/; fi
  fi
done
fi
du -sh gpurun_out
