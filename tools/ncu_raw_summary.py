"""Print the handful of `ncu --page raw --csv` metrics we track, per kernel.
Usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_elapsed.max', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get('Kernel Name'))
    for w in want:
        if w in d: print("  %-70s %s %s" % (w, d[w], units[hdr.index(w)]))
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            v = float(d[k])
            if v > 0.1: st.append((v, k.split('issue_stalled_')[1].split('_per')[0]))
    print("  stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)))
