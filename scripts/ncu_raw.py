"""Print the key metrics + stall reasons of an `ncu --page raw --csv` dump (first kernel row)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:3 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
    for k in keys:
        if k in d: print(f"{k:70s} {d[k][0]} {d[k][1]}")
    st = {k.split('stalled_')[1]: int(d[k][0]) for k in d if re.search(r'pcsamp_warps_issue_stalled_(?!.*not_issued)', k)}
    tot = sum(st.values())
    print('stalls:', ', '.join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v * 100 > tot))
