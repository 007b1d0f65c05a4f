"""Prints a compact table of the roofline / stall metrics of every kernel in an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('=====', r[ki][:70])
    for i, h in enumerate(hdr):
        if h in want:
            print('  %-86s %s %s' % (h, r[i], units[i]))
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    for v, h in sorted(st, reverse=True)[:5]:
        print('    stall %-60s %.2f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
