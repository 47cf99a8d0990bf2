#!/bin/bash
# multi-GPU evidence:  bash scratch/r2_multi.sh N   (run under gpurun --gpus N)
set -u
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "N=%d"%d["n_gpus"], d["scaling"], "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "hist/step %.3g seg/step %.3g"%(d["histories_per_step"], d["segments_per_step"]), d.get("tally_modes_run"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-600:])
PY
}
{
nvidia-smi -L | wc -l
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3; fi
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 5 2>&1 | tail -1 > gpurun_out/r2m_weak_f32_n$N.json; show gpurun_out/r2m_weak_f32_n$N.json
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --global-particles 1000000000 2>&1 | tail -1 > gpurun_out/r2m_strong_f32_n$N.json; show gpurun_out/r2m_strong_f32_n$N.json
if [ "$N" = "8" ]; then
  timeout 400 $TR bench.py --gpus $N --workload crookedpipe_f64 --steps 10 --warmup 5 2>&1 | tail -1 > gpurun_out/r2m_weak_f64_n$N.json; show gpurun_out/r2m_weak_f64_n$N.json
  timeout 400 $TR bench.py --gpus $N --workload crookedpipe_f64 --steps 10 --warmup 5 --scaling strong --global-particles 100000000 2>&1 | tail -1 > gpurun_out/r2m_strong_f64_n$N.json; show gpurun_out/r2m_strong_f64_n$N.json
fi
} 2>&1 | tee gpurun_out/r2_multi_n$N.log
