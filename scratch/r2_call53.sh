#!/bin/bash
# refill threshold (idle lanes before a warp refills) on the workloads with the lowest lane occupancy
set -u
mkdir -p gpurun_out
{
for wl in crookedpipe_f64 marshak_f32_rw; do for m in 2 4 8 12; do IMC_REFILL_MIN=$m timeout 200 python bench.py --workload $wl --track refill --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl refill_min $m: ms/step %.3f kernel %.3f frac %.4f'%(d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac']))"; done; done
} 2>&1 | tee gpurun_out/r2_call53.log
