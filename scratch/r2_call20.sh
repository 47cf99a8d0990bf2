#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
for wl in suolson_f16 crookedpipe_f64; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.2f kernel %.2f frac %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d.get('tally_modes_run'))"; done
} 2>&1 | tee gpurun_out/r2_call20.log
