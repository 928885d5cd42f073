#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ai; mkdir -p $O
timeout 200 python scripts/one_sort.py 28 > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_sort.csv python scripts/one_sort.py 28 > $O/ncu.log 2>&1
python - "$O/launches_sort.csv" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:64]), {})[r[mi]]=r[vi]
ids=sorted(d); half=len(ids)//2
tot=0
for k in ids[half:]:
    m=d[k]; t=float(m.get("gpu__time_duration.sum",0)); tot+=t
    if t>200000: print(k[0], k[1].ljust(64), int(t), m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"))
print("total ms of the second call", tot/1e6)
PY
