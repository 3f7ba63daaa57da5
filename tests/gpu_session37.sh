#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 nsys --version > /dev/null 2>&1 && echo nsys available
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s37_launches_c2.csv python tests/perf_probe.py c2clip > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/s37_launches_c2.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); idc=h.index('ID')
seq=[(int(r[idc]), r[kn].split('(')[0], float(r[mv].replace(',',''))) for r in rows[hi+2:] if len(r)>mv]
# last frame = last third
n=len(seq)//3
last=seq[-n:]
agg=collections.OrderedDict()
for _,k,v in last: agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print("%-22s n=%2d total %.3f ms"%(k,len(v),sum(v)/1e6))
PY
