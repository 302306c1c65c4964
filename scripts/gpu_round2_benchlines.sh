#!/bin/bash
# Bench lines of every BASELINE config on ONE GPU (final state), into gpurun_out/final/.
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_voxel_detr.err | tail -1 > $O/bench_voxel_detr.json
for w in conquer centerpoint_waymo centerpoint_nusc config1; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 2>$O/bench_$w.err | tail -1 > $O/bench_$w.json
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
