"""Dev tool: key metrics per captured launch of an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
 'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
want += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
idx = {h: i for i, h in enumerate(hdr)}
w = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for m in want:
    if m in idx:
        name = m.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
        print(f"{name} [{units[idx[m]]}]: " + ' | '.join(d[idx[m]][:w] for d in data))
