#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench (driver-like)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2f2_bench20.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2f2_bench20.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g (%.2f ms/step)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["schedule_per_step"], d["clocks"], "launches/step", d["gpu_launches"]/d["steps"])
PY
for wl in crookedpipe_f64 marshak_f32_rw suolson_f32 suolson_f64 suolson_f16; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.3f kernel %.3f frac %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d.get('tally_modes_run'), d['schedule_per_step'][-1])"; done
} 2>&1 | tee gpurun_out/r2_final2.log
