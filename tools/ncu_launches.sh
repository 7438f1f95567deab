#!/bin/bash
# Per-kernel launch list (gpu__time_duration.sum) of one bench configuration, summarised per kernel name.
# Usage: tools/ncu_launches.sh <tag> <bench args...>
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
python bench.py --no-cpu "$@" > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --no-cpu "$@" > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py launches $OUT/launches_$TAG.csv > $OUT/launches_$TAG.txt 2>&1 || python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv', errors='replace')) if len(r)>5 and r[0].isdigit()]
hdr=None
for r in csv.reader(open('$OUT/launches_$TAG.csv', errors='replace')):
    if r and r[0]=='ID': hdr=r; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows:
    v=float(r[vi].replace(',','')); u=r[ui]
    us = v/1000.0 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1000.0)
    a=agg.setdefault(r[ki],[0,0.0]); a[0]+=1; a[1]+=us
tot=sum(a[1] for a in agg.values())
with open('$OUT/launches_$TAG.txt','w') as f:
    for k,(n,us) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write('%-70s %5d %12.1f us %9.1f us/launch %5.1f%%\n'%(k[:70],n,us,us/n,100*us/tot))
    f.write('total %.1f us\n'%tot)
PY
head -40 $OUT/launches_$TAG.txt
