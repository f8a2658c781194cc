"""Print the handful of `ncu --page raw --csv` metrics the DESIGN/roofline discussion uses, for every kernel in a report."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('wait', 'barrier', 'short_scoreboard', 'long_scoreboard', 'math_pipe_throttle', 'not_selected', 'no_instruction', 'mio_throttle',
         'branch_resolving', 'dispatch_stall', 'lg_throttle')]
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get('Kernel Name', '?')[:100], " grid", d.get('Grid Size'), " block", d.get('Block Size'))
    for h, u in zip(hdr, units):
        if h in WANT:
            print("  %-88s %-16s %s" % (h, u, d[h]))
