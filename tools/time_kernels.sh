#!/bin/bash
# usage: time_kernels.sh tag [ENV=VAL...] : per-kernel gpu time (ncu, serialised) of one step of the default bench
tag=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/tk_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/tk_$tag.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ik].split('(')[0][:50]
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[iv].replace(',',''))
print("$tag", {k:(v[0], round(v[1]/1e6/v[0],3)) for k,v in agg.items() if v[1]>2e5})
PY
