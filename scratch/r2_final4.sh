#!/bin/bash
# re-validation of HEAD after the census-tally refactor (imc_warp_runs.cuh)
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in suolson_f32 suolson_f64; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.3f kernel %.3f frac %.4f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac']))"; done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2h_bench20.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench20.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g (%.2f ms/step)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["clocks"])
PY
} 2>&1 | tee gpurun_out/r2_final4.log
