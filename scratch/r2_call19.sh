#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== suolson_f16 launch list (skip reducer default on)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c19_f16_launches.csv python bench.py --workload suolson_f16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c19_f16.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c19_f16_launches.csv")) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
import re
last=collections.OrderedDict()
ids=[int(r[0]) for r in rows]
n=len(rows)
# take the last quarter of launches (the last timed steps)
tail=rows[int(n*0.8):]
agg=collections.defaultdict(lambda:[0,0.0])
for r in tail:
    name=re.sub(r"\(.*","",r[4])[:70]; v=float(r[-1].replace(",",""))
    unit=r[-2]
    if unit=="ns": v/=1e6
    elif unit=="us" or unit=="usecond": v/=1e3
    elif unit=="second": v*=1e3
    agg[name][0]+=1; agg[name][1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{t:9.3f} ms {c:4d}x {k}")
print("unit of last row:", tail[-1][-2])
PY
echo "== suolson_f16 / f32 bench"
for wl in suolson_f16 suolson_f32; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl value %.4g ms/step %.2f kernel %.2f frac %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['tracking_kernel_ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d.get('tally_modes_run'))"; done
} 2>&1 | tee gpurun_out/r2_call19.log
