#!/bin/bash
# ncu launch list of a bench run:  bash scratch/launch_list.sh TAG [bench args]
set -u
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_launches.log 2>&1
python - gpurun_out/${TAG}_launches.csv <<'PY'
import csv, collections, re, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
def ms(r):
    v=float(r[-1].replace(",","")); u=r[-2]
    return v/1e6 if u in ("ns","nsecond") else v/1e3 if u in ("us","usecond") else v*1e3 if u in ("s","second") else v
# the resident timed steps: launches between the 3rd and 5th tracking kernel
trk=[i for i,r in enumerate(rows) if "k_track" in r[4]]
print("launches", len(rows), "tracking launches", len(trk))
lo, hi = trk[3] if len(trk)>4 else 0, trk[4] if len(trk)>4 else len(rows)
# one step = from one tracking launch to the next
step=rows[lo:hi]
agg=collections.OrderedDict()
for r in step:
    name=re.sub(r"\(.*","",r[4]); name=re.sub(r"^void ","",name)[:60]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms(r)
tot=sum(v[1] for v in agg.values())
print("one step: %d launches, %.3f ms of kernels"%(len(step),tot))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{t:9.3f} ms {100*t/tot:5.1f}% {c:3d}x {k}")
PY
