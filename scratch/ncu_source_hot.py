"""Hot SASS lines of one kernel from an ncu report: python scratch/ncu_source_hot.py rep.ncu-rep kernel_regex [section_index]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = []
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        secs.append([r[1], rows[i + 1], []])
    elif secs and len(r) > 10 and r[0] != 'Address':
        secs[-1][2].append(r)
print(len(secs), 'sections:', [s[0][:40] for s in secs])
name, hdr, data = secs[which]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ts = sum(int(r[isamp]) for r in data); te = sum(int(r[iex]) for r in data)
print(name[:80], 'lines', len(data), 'samples', ts, 'warp-instr', te)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:45]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {100*int(r[isamp])/ts:5.1f}% ex {int(r[iex]):9d} {r[ia].strip()[:100]}")
for i in range(0, len(data), 100):
    e = sum(int(r[iex]) for r in data[i:i+100]); s = sum(int(r[isamp]) for r in data[i:i+100])
    print(f"lines {i:5d}-{i+99:5d}: exec {100*e/te:5.1f}% samples {100*s/ts:5.1f}%")
