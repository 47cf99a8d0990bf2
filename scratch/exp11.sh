#!/bin/bash
# where do the 85 ms go?  timing-only builds (wrong physics): 1 = no deposit RED, 3 = cell gathers confined to 8 KB, 4 = both
run() { python bench.py --track refill --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['tracking_kernel_ms_per_step']; print('   seg/s %.4g  kernel_ms %.2f  seg/hist %.1f  ns/1e3seg %.3f' % (d['value'], s, d['segments_per_history'], 1e6*s/(d['value']*d['ms_per_step']*1e-3)*1e3))"; }
for cfg in "-DIMC_DEBUG_TALLY=0" "-DIMC_DEBUG_TALLY=1" "-DIMC_DEBUG_TALLY=3" "-DIMC_DEBUG_TALLY=4"; do
  IMC_NVCC_EXTRA="$cfg" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
  echo "== $cfg"; run
done
