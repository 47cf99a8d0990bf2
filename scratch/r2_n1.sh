#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "N=%d"%d["n_gpus"], d["scaling"], "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["tracking_kernel_ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], "hist/step %.3g seg/step %.3g"%(d["histories_per_step"], d["segments_per_step"]), d.get("tally_modes_run"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-600:])
PY
}
{
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_libm_policy.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2m_weak_f32_n1.json; show gpurun_out/r2m_weak_f32_n1.json
timeout 400 python bench.py --workload crookedpipe_f64 --steps 10 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2m_weak_f64_n1.json; show gpurun_out/r2m_weak_f64_n1.json
timeout 900 python bench.py --steps 5 --warmup 3 --scaling strong --global-particles 1000000000 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2m_strong_f32_n1.json; show gpurun_out/r2m_strong_f32_n1.json
} 2>&1 | tee gpurun_out/r2_n1.log
