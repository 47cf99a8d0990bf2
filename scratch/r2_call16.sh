#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1].split("/")[-1], "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g"%d["e2e"]["value"], d["schedule_per_step"][-1], d.get("tally_modes_run"), "launches/step %.0f"%(d["gpu_launches"]/d["steps"]))
PY
}
{
echo "== whole GPU suite"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== crookedpipe_f32 PAIRWISE TRUE (AUTO -> FIXED at this scale) vs FALSE"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pairwise TRUE 2>&1 | tail -1 > gpurun_out/r2c16_pw_true.json; show gpurun_out/r2c16_pw_true.json
for wl in crookedpipe_f64 marshak_f32_rw suolson_f32 suolson_f64 suolson_f16; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c16_$wl.json; show gpurun_out/r2c16_$wl.json
done
} 2>&1 | tee gpurun_out/r2_call16.log
