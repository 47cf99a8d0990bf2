#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== driver-like bench: 20 steps after 5 warm-up"
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2f_bench20.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench20.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g (%.2f ms/step)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["schedule_per_step"][-1], d["clocks"], "launches/step", d["gpu_launches"]/d["steps"], "cpu", d["cpu_baseline"]["value"], d["tally_rel_err"])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2f_ref.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_ref.json"))
print("ref value %.4g"%d["value"], d["cpu_baseline"]["cores"], "single %.4g"%d["cpu_baseline"]["single_core_value"], d["config"]["workload"])
PY
bash scratch/ncu_refill.sh r2f
bash scratch/launch_list.sh r2f_f32 | head -30
} 2>&1 | tee gpurun_out/r2_final1.log
