#!/bin/bash
# Round 2, GPU call 1 (1 GPU): re-validate HEAD of round 1 and run what round 1 left unmeasured.
set -u
mkdir -p gpurun_out
{
nvidia-smi -L
echo "== 1. opt-in tests (stagnation-skip reducer, recorded-run replay on the GPU)"
IMC_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests -m gpu -q -k "stagnation or replays_recorded" 2>&1 | tail -5
echo "== 1b. suolson_f16 (sequential Float16 EXACT sums) without / with IMC_EXACT_SKIP"
timeout 200 python bench.py --workload suolson_f16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=0', d['ms_per_step'], d['value'])"
IMC_EXACT_SKIP=1 timeout 200 python bench.py --workload suolson_f16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=1', d['ms_per_step'], d['value'])"
echo "== 2. whole GPU suite"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 3. bench lines, N = 1"
for wl in crookedpipe_f32 crookedpipe_f64 marshak_f32_rw suolson_f32 suolson_f64; do
  timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c1_bench_$wl.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2c1_bench_$wl.json"))
print("$wl", "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g"%d["e2e"]["value"], d["schedule_per_step"][-1], d["clocks"])
PY
done
} 2>&1 | tee gpurun_out/r2_call1.log
