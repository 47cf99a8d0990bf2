#!/bin/bash
run() { python bench.py --track refill --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   seg/s %.4g  ms/step %.1f  kernel_ms %.2f' % (d['value'], d['ms_per_step'], d['tracking_kernel_ms_per_step']))"; }
for cfg in "-DIMC_TRACK_THREADS=192 -DIMC_TRACK_MIN_BLOCKS=6" "-DIMC_TRACK_THREADS=192 -DIMC_TRACK_MIN_BLOCKS=6 -DIMC_EARLY_NEXT_CELL=0" "-DIMC_TRACK_THREADS=128 -DIMC_TRACK_MIN_BLOCKS=9 -DIMC_EARLY_NEXT_CELL=0 -DIMC_FASTDIV_DIR=0" "-DIMC_TRACK_THREADS=224 -DIMC_TRACK_MIN_BLOCKS=5 -DIMC_EARLY_NEXT_CELL=0 -DIMC_FASTDIV_DIR=0" "-DIMC_TRACK_THREADS=96 -DIMC_TRACK_MIN_BLOCKS=12 -DIMC_EARLY_NEXT_CELL=0 -DIMC_FASTDIV_DIR=0"; do
  IMC_NVCC_EXTRA="$cfg" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
  echo "== $cfg"; cuobjdump --dump-resource-usage build/imc_engine_f32.o 2>/dev/null | grep -A1 "k_track_refillINS_3F32ELi2ELb0ELi0" | grep -oE "REG:[0-9]+ STACK:[0-9]+"; run
done
