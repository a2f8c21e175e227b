"""Summarise an .ncu-rep (raw + source pages) -> dict; used to build profiles/*.json"""
import csv, json, subprocess, sys
def raw(rep, index=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + index]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
def summary(rep, index=0):
    d = raw(rep, index)
    out = {'kernel': d['Kernel Name'][1]}
    out.update({k: {'unit': d[k][0], 'value': d[k][1]} for k in KEYS if k in d})
    stalls = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(v) for h, (u, v) in d.items()
              if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h and v.isdigit() and int(v) > 0}
    tot = sum(stalls.values())
    out['stall_samples_pct'] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])}
    return out
def opcodes(rep, top=12, kernel=None):
    cmd = ['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--kernel-name', 'regex:' + kernel] if kernel else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr) or not (r[ix['# Samples']] or '0').isdigit(): continue
        src = r[ix['Source']]; toks = src.split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        a = agg.setdefault(op, [0, 0]); a[0] += int(r[ix['# Samples']] or 0); a[1] += int(r[ix['Instructions Executed']] or 0)
    return {op: {'samples': a[0], 'inst_executed': a[1]} for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]}
if __name__ == '__main__':
    rep = sys.argv[1]
    index = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kernel = sys.argv[3] if len(sys.argv) > 3 else None
    s = summary(rep, index); s['opcodes'] = opcodes(rep, kernel=kernel)
    print(json.dumps(s, indent=1))
