#!/bin/bash
mkdir -p gpurun_out
python tools/c4_pass1_breakdown.py > gpurun_out/r22_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r22_c4_launches.csv python tools/c4_pass1_breakdown.py 1e9 >> gpurun_out/r22_c4.txt 2>&1
cat gpurun_out/r22_c4.txt
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r22_c4_launches.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",",""));
    if r[ui]=="ns": v/=1e6
    elif r[ui]=="us": v/=1e3
    elif r[ui]=="s": v*=1e3
    k=r[ki][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{t:9.3f} ms {n:5d}x {k}")
PY
