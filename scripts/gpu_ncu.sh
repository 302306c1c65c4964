#!/bin/bash
# ncu evidence: (1) launch list with per-kernel device time for one bench invocation,
# (2) one --set full capture of the kernel family given as $1 (regex), default the sparse conv GEMM.
mkdir -p gpurun_out
PAT=${1:-spconv}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:$PAT -s ${SKIP:-40} -c ${COUNT:-3} -f -o gpurun_out/prof_$PAT \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
