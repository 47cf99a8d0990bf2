#!/bin/bash
# full GPU suite at HEAD (clean() takes the survivors from the transport's census count) + step times of the small-step decks
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
for wl in suolson_f32 suolson_f64 marshak_f32_rw crookedpipe_f64; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.3f kernel %.3f frac %.4f e2e %.4g launches/step %.1f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches']/d['steps']), d.get('tally_modes_run'), d['schedule_per_step'][-1])"; done
timeout 300 python scratch/shipped_decks.py 2>&1 | tail -8
} 2>&1 | tee gpurun_out/r2_call49.log
