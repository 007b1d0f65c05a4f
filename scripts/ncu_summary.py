"""Prints the key roofline / stall metrics of an .ncu-rep (first kernel matching a substring). Dev + profiles/ helper."""
import csv, subprocess, sys
rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
# a .csv argument is the output of `ncu -i X.ncu-rep --page raw --csv` (what scripts/gpu_round.sh brings back from the GPU box)
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # n-th launch matching the pattern
hits = [r for r in rows[2:] if pat in r[ki]]
r = hits[min(nth, len(hits) - 1)]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg']
print(r[ki][:100])
for i, h in enumerate(hdr):
    if h in want or h.startswith('smsp__average_warp') or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')) or 'tensor' in h:
        print('  %-90s %s %s' % (h, r[i], units[i]))
