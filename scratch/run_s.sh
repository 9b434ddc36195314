#!/bin/bash
python scratch/solve_prof.py 4097
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve.csv python scratch/solve_prof.py 4097 > gpurun_out/solve_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_solve.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg={}
for r in rows[1:]:
    k=r[ki][:90]; v=float(r[vi].replace(',',''))
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print("total kernel time ms", tot/1e6, "launches", sum(a[0] for a in agg.values()))
for k,(n,t) in sorted(agg.items(),key=lambda x:-x[1][1])[:25]: print(f"{n:6d} {t/1e6:9.3f} ms  {t/n/1000:9.2f} us avg  {k}")
PY
