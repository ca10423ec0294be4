#!/bin/bash
# usage: tools/gpu_launches.sh TAG [shape] -> ncu launch list (durations only) of the device-resident steps
TAG=$1; SHAPE=${2:-pacbio}
python bench.py --no-cpu-baseline --text-sample 0 --steps 2 --warmup 3 --resident-only --shape $SHAPE --parity-reads 0 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --no-cpu-baseline --text-sample 0 --steps 2 --warmup 3 --resident-only --shape $SHAPE --parity-reads 0 > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches_$TAG.csv") if not l.startswith("=="))]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg={}
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki].split("(")[0][:40]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("%-42s %4d %10.1f us/launch %5.1f%%"%(k,a[0],a[1]/a[0]/1e3,100*a[1]/tot))
PY
