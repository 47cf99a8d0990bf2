#!/bin/bash
# A/B of build variants on one box:  bash scratch/ab.sh name1 name2[:ENV=val[,ENV2=val2]] ...
# (variants/libimc_<name>.so; two rounds each, interleaved; BENCH_ARGS adds bench.py flags; AB_TAG names the log)
set -u
mkdir -p gpurun_out
run() {
  local spec=$1 name=${1%%:*} envs=""
  [[ "$spec" == *:* ]] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  env $envs IMC_LIB=$PWD/variants/libimc_$name.so python bench.py --track refill --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$spec: seg/s %.4g  ms/step %.2f  kernel_ms %.2f  frac %.4f  sph %.1f' % (d['value'], d['ms_per_step'], d['tracking_kernel_ms_per_step'], d['roofline']['frac'], d['segments_per_history']))"; }
for round in 1 2; do for v in "$@"; do run $v; done; done 2>&1 | tee -a gpurun_out/ab_${AB_TAG:-x}.log
