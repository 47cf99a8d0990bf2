#!/bin/bash
run() { python bench.py --track refill --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   seg/s %.4g  ms/step %.1f  kernel_ms %.2f' % (d['value'], d['ms_per_step'], d['tracking_kernel_ms_per_step']))"; }
IMC_NVCC_EXTRA="" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
echo "== default, tally fixed"; run --tally fixed
for cfg in "-DIMC_DEBUG_TALLY=1" "-DIMC_DEBUG_TALLY=2"; do
  IMC_NVCC_EXTRA="$cfg" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
  echo "== $cfg"; run
done
