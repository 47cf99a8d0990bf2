#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_checkpoint.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
AB_TAG=c15 bash scratch/ab.sh q4 q5 q6
echo "== default bench with the checkpointed e2e loop"
timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c15_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c15_bench.json"))
print("value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g ms/step %.2f"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), "seg/step", d["segments_per_step"], d["e2e"]["segments_per_step"], "traffic", d["roofline"]["traffic"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("tally_rel_err"))
PY
} 2>&1 | tee gpurun_out/r2_call15.log
