"""Samples / executed instructions of one kernel between consecutive BAR.SYNC instructions, plus the hottest SASS lines.
python scratch/ncu_regions.py rep.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        hdr = rows[i + 1]; data = [x for x in rows[i + 2:] if len(x) > 10 and x[0] != 'Address']; break
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ist = hdr.index('Warp Stall Sampling (All Samples)')
ts = sum(int(r[isamp]) for r in data); te = sum(int(r[iex]) for r in data)
print('lines', len(data), 'samples', ts, 'warp-instr', te)
bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[ia]]
prev = 0
for b in bars + [len(data)]:
    e = sum(int(r[iex]) for r in data[prev:b]); s = sum(int(r[isamp]) for r in data[prev:b])
    if s / ts > 0.004:
        ops = {}
        for r in data[prev:b]:
            op = r[ia].split()[0] if not r[ia].strip().startswith('@') else r[ia].split()[1]
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + int(r[iex])
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
        print(f"lines {prev:5d}-{b:5d}: exec {100*e/te:5.1f}% samples {100*s/ts:5.1f}%  n={b-prev}  " + ' '.join(f"{k}:{100*v/max(e,1):.0f}%" for k, v in top))
    prev = b
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {100*int(r[isamp])/ts:5.1f}% ex {int(r[iex]):9d} {r[ia].strip()[:90]}")
