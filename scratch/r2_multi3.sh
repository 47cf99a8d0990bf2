#!/bin/bash
# two GPUs at HEAD: the driver's launch of bench.py, and the Su-Olson deck (lazy compaction on every rank)
set -u
mkdir -p gpurun_out
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2g_n2.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_n2.json"))
print("N=2 crookedpipe_f32 value %.4g ms/step %.2f kernel %s e2e %.4g"%(d["value"], d["ms_per_step"], d["tracking_kernel_ms_per_step_by_rank"], d["e2e"]["value"]), d["clocks"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload suolson_f32 --steps 8 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2g_n2_su32.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_n2_su32.json"))
print("N=2 suolson_f32 value %.4g ms/step %.2f kernel %s e2e %.4g resident %s"%(d["value"], d["ms_per_step"], d["tracking_kernel_ms_per_step_by_rank"], d["e2e"]["value"], d["particles_resident"]))
PY
} 2>&1 | tee gpurun_out/r2_multi3.log
