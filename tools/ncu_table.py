"""ncu -i rep --page raw --csv | table of selected metrics for every captured launch"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size']
extra = sys.argv[2:]
for k in keys + extra:
    if k in ix:
        print(f"{k[:70]:70s} [{rows[1][ix[k]]}]", *[r[ix[k]][:22] for r in rows[2:]], sep=' | ')
if '--stalls' in sys.argv:
    for h in hdr:
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h:
            print(f"{h[33:]:40s}", *[r[ix[h]] for r in rows[2:]], sep=' | ')
