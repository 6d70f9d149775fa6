import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__grid_size','sm__cycles_elapsed.avg','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__maximum_warps_per_active_cycle_pct']
stalls=[h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')] or [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
idx={h:i for i,h in enumerate(hdr)}
seen=set()
for r in data:
    name=r[idx['Kernel Name']].split('(')[0]
    if name in seen: continue
    seen.add(name)
    print('====',name)
    for w in want:
        if w in idx: print(f'  {w:68s} {r[idx[w]]:>18s} {units[idx[w]]}')
    st=[(float(r[idx[h]].replace(',','')),h) for h in stalls if r[idx[h]] not in ('','n/a')]
    for v,h in sorted(st,reverse=True)[:8]:
        print(f'     stall {h.split("issue_stalled_")[1][:40]:40s} {v:8.2f}')
