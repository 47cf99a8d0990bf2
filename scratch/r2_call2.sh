#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): bit-identity 1 vs 2 GPUs at HEAD (after the stream-ordering fix 9597cc9) + the N = 2 bench line
set -u
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c2_bench_n2.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c2_bench_n2.json"))
print("N=2", "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["clocks"])
PY
} 2>&1 | tee gpurun_out/r2_call2.log
