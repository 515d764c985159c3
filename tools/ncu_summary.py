"""Print key metrics from an .ncu-rep (first profiled kernel) and optionally the hottest source lines."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__cycles_active.avg', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_issue_stalled_short_scoreboard.pct',
        'smsp__inst_executed_op_shfl.sum' ]
for k in want:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:70s} {units[i]:12s} {vals[i]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        try:
            if float(vals[i]) > 3: print(f"{h:90s} {vals[i]}")
        except: pass
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    for hi, r in enumerate(rows):
        if 'Source' in r and any('Samples' in c for c in r): break
    h = rows[hi]
    si = h.index('Source'); 
    ci = [i for i,c in enumerate(h) if c.strip() in ('# Samples','Warp Stall Sampling (All Samples)','Samples')]
    ii = [i for i,c in enumerate(h) if c.strip() == 'Instructions Executed']
    print(h[:12])
    agg = []
    for r in rows[hi+1:]:
        try: agg.append((int(r[ci[0]]), int(r[ii[0]]) if ii else 0, r[0], r[si][:110]))
        except: pass
    tot = sum(a[0] for a in agg); toti = sum(a[1] for a in agg)
    print('total samples', tot, 'total inst', toti)
    for a in sorted(agg, reverse=True)[:int(sys.argv[2])]:
        print(f"{100*a[0]/max(tot,1):5.1f}% smp {100*a[1]/max(toti,1):5.1f}% ins  L{a[2]:>5s} {a[3]}")
