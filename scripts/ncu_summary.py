"""Key raw metrics of the first kernel in an ncu report (scratch tool).  usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, v = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
for i, n in enumerate(h):
    if n in want or ('issue_stalled' in n and 'per_issue_active' in n and 'not_issued' not in n and float(v[i] or 0) > 0.05):
        print(f"{n:85s} {v[i]}")
