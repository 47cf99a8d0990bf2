#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out.
#   gpurun --timeout 900 -- 'bash scratch/r2_first_call.sh'          (1 GPU; items 2b / 2c need --gpus 2)
set -u
mkdir -p gpurun_out
{
echo "== 1. opt-in tests (stagnation-skip reducer, recorded-run replay on the GPU)"
IMC_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests -m gpu -q -k "stagnation or replays_recorded" 2>&1 | tail -5
echo "== 1b. suolson_f16 (sequential Float16 EXACT sums) without / with IMC_EXACT_SKIP"
timeout 200 python bench.py --workload suolson_f16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=0', d['ms_per_step'], d['value'])"
IMC_EXACT_SKIP=1 timeout 200 python bench.py --workload suolson_f16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=1', d['ms_per_step'], d['value'])"
echo "== 2. whole GPU suite and the default bench line on this tree"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python bench.py 2>&1 | tail -1 | cut -c1-400
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== 2b. two GPUs: bit-identity test and the bench line after the ordering fix of dist.py"
  timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
fi
} 2>&1 | tee gpurun_out/r2_first_call.log
