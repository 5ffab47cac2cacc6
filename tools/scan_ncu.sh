#!/bin/bash
# launch list (device time per kernel) of the last of four scans of tools/scan_probe.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/scan_launches.csv python tools/scan_probe.py ${1:-20000} ${2:-5000} > gpurun_out/scan_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/scan_launches.csv")) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
names=[(r[4], r[6], r[7], r[8], float(r[-1].replace(",",""))) for r in rows]
n=len(names)
# the probe runs 4 scans; take the last quarter of the launches after the uploads
last=names[-(n//4):] if n>=8 else names
agg=collections.OrderedDict()
for k,st,bs,gs,t in last:
    k=re.sub(r"\(.*","",k)
    a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=t; a[2]=max(a[2],t)
tot=sum(a[1] for a in agg.values())
print("last scan: %d launches, %.2f ms of serialized kernel time"%(len(last),tot/1e6))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print("%8.1f us total  %8.1f us max  x%-4d %s"%(a[1]/1e3,a[2]/1e3,a[0],k[:110]))
PY
