"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ cite."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum ', 'sm__pipe_tensor', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.avg ',
        'smsp__inst_executed_op_shared', 'sm__sass_inst_executed_op_shared', 'smsp__sass_thread_inst_executed_op_ffma',
        'smsp__sass_thread_inst_executed_op_fmul', 'smsp__sass_thread_inst_executed_op_fadd', 'lts__throughput.avg.pct']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('== kernel:', name[:100])
    for h, u, v in zip(hdr, units, vals):
        hh = h + ' '
        if any(k in hh for k in KEYS) and v not in ('', '0'):
            print(f'  {h:88s} {u:14s} {v}')
