#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
for b in 16 64 256; do IMC_EVENT_BATCH=$b timeout 300 python bench.py --track event --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('event batch $b: value %.4g ms/step %.2f kernel %.2f frac %.4f launches/step %.0f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'], d['gpu_launches']/d['steps']), d['schedule_per_step'][-1])"; done
for wl in suolson_f32 crookedpipe_f64; do IMC_EVENT_BATCH=64 timeout 300 python bench.py --workload $wl --track event --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl event 64: value %.4g ms/step %.2f kernel %.2f frac %.4f launches/step %.0f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'], d['gpu_launches']/d['steps']), d['schedule_per_step'][-1])"; done
} 2>&1 | tee gpurun_out/r2_call36.log
