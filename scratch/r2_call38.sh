#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
for wl in suolson_f16 suolson_f32; do for pw in FALSE TRUE; do timeout 300 python bench.py --workload $wl --pairwise $pw --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl pairwise $pw value %.4g ms/step %.3f kernel %.3f frac %.4f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac']), d.get('tally_modes_run'), d['schedule_per_step'][-1])"; done; done
} 2>&1 | tee gpurun_out/r2_call38.log
