#!/bin/bash
# launch list (device time of every kernel) of a short bench run.  usage: gpu_launches.sh <tag> <skip> <count> <bench args...>
tag=$1; skip=$2; cnt=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c $cnt --csv --log-file gpurun_out/${tag}_launches.csv python bench.py "$@" > gpurun_out/${tag}_launches.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[iv].replace(',','')); u=r[iu]
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    k=r[ik].split('(')[0][-40:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:42s} n={n:4d} total {t:10.1f} us  avg {t/n:9.1f} us  {100*t/tot:5.1f}%")
PY
