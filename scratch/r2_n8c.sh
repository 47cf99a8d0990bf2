#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1].split("/")[-1], "N=%d"%d["n_gpus"], d["scaling"], "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "e2e %.4g (%.2f)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), ["%.1f"%x for x in d["tracking_kernel_ms_per_step_by_rank"]], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2f_weak_f32_n8.json; show gpurun_out/r2f_weak_f32_n8.json
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --global-particles 1000000000 2>&1 | tail -1 > gpurun_out/r2f_strong_f32_n8.json; show gpurun_out/r2f_strong_f32_n8.json
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-300
