#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2af; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py tests/test_gpu_parity.py tests/test_gpu_extras.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted or cum or sort or narrowed" > $O/pytest_scan.log 2>&1
tail -3 $O/pytest_scan.log | cut -c1-300
timeout 600 python bench.py --config C4 --steps 5 --warmup 3 > $O/bench_C4.json 2> $O/bench_C4.err
python - "$O/bench_C4.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
[print("   ", p) for p in d.get("points", [])]
PY
timeout 200 python scripts/one_unsorted.py 28 > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/launches_unsorted.csv python scripts/one_unsorted.py 28 > $O/ncu.log 2>&1
python - "$O/launches_unsorted.csv" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:60]), {})[r[mi]]=r[vi]
ids=sorted(d); half=len(ids)//2
for k in ids[half:]:
    m=d[k]
    if float(m.get("gpu__time_duration.sum",0))>100000: print(k[0], k[1].ljust(60), m.get("gpu__time_duration.sum"), m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"))
PY
