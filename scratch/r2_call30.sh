#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_libm_policy.py tests/test_checkpoint.py -m gpu -x -q 2>&1 | tail -4
for wl in marshak_f32_rw suolson_f32; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.3f kernel %.3f frac %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d.get('tally_modes_run'), d['schedule_per_step'][-1])"; done
} 2>&1 | tee gpurun_out/r2_call30.log
