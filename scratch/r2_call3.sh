#!/bin/bash
# Round 2, GPU call 3: per-thread counters + Float32 deposit REDs: tests, bench, one full ncu capture of the refill kernel with source
set -u
mkdir -p gpurun_out
{
echo "== GPU suite"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench (refill)"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c3_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c3_bench.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g"%d["e2e"]["value"], d["schedule_per_step"][-1])
PY
echo "== ncu full, refill kernel, 1 launch after 4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_track_refill -s 4 -c 1 -o gpurun_out/r2c3_refill -f python bench.py --track refill --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_ncu.log 2>&1
tail -3 gpurun_out/r2c3_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
} 2>&1 | tee gpurun_out/r2_call3.log
