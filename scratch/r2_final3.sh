#!/bin/bash
# end-of-round validation of HEAD on one B200: full GPU suite, smoke, the driver's bench command and its reference arm, the ncu
# capture + launch list of the headline workload, the other BASELINE configs
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (driver-like)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2g_bench20.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench20.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "sph %.1f"%d["segments_per_history"], "e2e %.4g (%.2f ms/step)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["schedule_per_step"][-1], d["clocks"], "launches/step", d["gpu_launches"]/d["steps"], "cpu", d["cpu_baseline"]["value"], d["tally_rel_err"])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2g_ref.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_ref.json"))
print("ref value %.4g"%d["value"], d["cpu_baseline"]["cores"], "single %.4g"%d["cpu_baseline"].get("single_core_value", 0), d["config"]["workload"])
PY
bash scratch/ncu_refill.sh r2g
bash scratch/launch_list.sh r2g_f32 | head -30
for wl in crookedpipe_f64 marshak_f32_rw suolson_f32 suolson_f64 suolson_f16; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.3f kernel %.3f frac %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d.get('tally_modes_run'), d['schedule_per_step'][-1])"; done
timeout 300 python bench.py --pairwise TRUE --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('crookedpipe_f32 PAIRWISE TRUE value %.4g ms/step %.3f kernel %.3f frac %.4f'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac']), d.get('tally_modes_run'))"
} 2>&1 | tee gpurun_out/r2_final3.log
