#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
( for i in 1 2 3 4 5 6 7 8 9 10 11 12; do sleep 3; nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader | tr '\n' ';'; echo; done > gpurun_out/r2_n8b_smi.log ) &
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 5 2>&1 | tail -1 > gpurun_out/r2m_weak_f32_n8b.json
wait
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m_weak_f32_n8b.json"))
print("N=8 value %.4g ms/step %.2f kernel(max) %.2f e2e %.4g"%(d["value"], d["ms_per_step"], d["tracking_kernel_ms_per_step"], d["e2e"]["value"]))
print("per rank kernel ms:", ["%.2f"%x for x in d["tracking_kernel_ms_per_step_by_rank"]])
PY
tail -4 gpurun_out/r2_n8b_smi.log
