"""Key numbers of one-kernel ncu reports:  python tools/ncu_brief.py file.ncu-rep [...]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shared_ld.sum']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', f, vals[hdr.index('Kernel Name')][:90])
        for i, h in enumerate(hdr):
            stall = 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')
            if h in WANT or (stall and float(vals[i].replace(',', '') or 0) > 0.05):
                print(f'   {h:84s} {vals[i]} {units[i]}')
