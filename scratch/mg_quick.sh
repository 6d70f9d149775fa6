#!/bin/bash
# quick GPU loop for the multigrid kernels: parity tests, 512^3 timing, per-kernel times
timeout 300 python -m pytest tests/test_gpu_mg.py -x -q 2>&1 | tail -2
timeout 200 python scratch/mg_check.py 512 2>&1 | grep -v jacobi | head -2
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_mg_q.csv python scratch/prof_mg.py 512 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_mg_q.csv', errors='ignore')))
for i, r in enumerate(rows):
    if 'Kernel Name' in r: hdr = r; start = i + 1; break
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi: continue
    n = r[ki].split('(')[0]; v = float(r[vi].replace(',', ''))
    a = agg.setdefault(n, {}); a.setdefault(r[mi], []).append(v)
for n, a in agg.items():
    t = a.get('gpu__time_duration.sum', [0]); ins = a.get('smsp__inst_executed.sum', [0])
    if max(t) > 20000: print(f"{n:34s} n={len(t):3d} max {max(t)/1e3:9.1f} us  inst {max(ins)/1e6:8.1f} M  sum {sum(t)/1e3:9.1f} us")
PY
