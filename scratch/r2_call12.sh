#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== driver-like bench: 20 steps after 5 warm-up"
timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2c12_bench20.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c12_bench20.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g"%d["e2e"]["value"], d["schedule_per_step"][-1], d["clocks"])
PY
bash scratch/ncu_refill.sh r2c12
} 2>&1 | tee gpurun_out/r2_call12.log
