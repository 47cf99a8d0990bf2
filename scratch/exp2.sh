#!/bin/bash
run() { python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   seg/s %.4g  ms/step %.1f  kernel_ms %.1f frac %.3f sched %s' % (d['value'], d['ms_per_step'], d['tracking_kernel_ms_per_step'], d['roofline']['frac'], d['schedule_per_step']))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for mb in 4 5; do
  IMC_NVCC_EXTRA="-DIMC_TRACK_MIN_BLOCKS=$mb" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
  echo "== MIN_BLOCKS=$mb crookedpipe_f32 refill"; run --track refill
  echo "== MIN_BLOCKS=$mb crookedpipe_f32 static"; run --track history
done
for mb in 3 4; do
  IMC_NVCC_EXTRA="-DIMC_TRACK_MIN_BLOCKS=$mb" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
  echo "== MIN_BLOCKS=$mb crookedpipe_f64 (1024^2, 1e8) auto"; run --workload crookedpipe_f64
done
IMC_NVCC_EXTRA="" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1
echo "== suolson_f32 auto"; run --workload suolson_f32
echo "== marshak_f32_rw"; run --workload marshak_f32_rw
