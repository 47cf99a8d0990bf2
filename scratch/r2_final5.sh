#!/bin/bash
# last validation of the round: energycheck served from the tally's readback
set -u
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 100 python bench.py --workload marshak_f32_rw --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('marshak_f32_rw value %.4g ms/step %.3f kernel %.3f launches/step %.1f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['gpu_launches']/d['steps']))"
} 2>&1 | tee gpurun_out/r2_final5.log
