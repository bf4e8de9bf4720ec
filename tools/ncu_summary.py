"""Summarise an .ncu-rep (raw page + source page) into the few numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--ops]"""
import csv, collections, re, subprocess, sys, io

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_tensor']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('==', name[:100])
    for h, u, v in zip(hdr, units, vals):
        if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.15):
            print(f'  {h:88s} {u:16s} {v}')
if '--ops' in sys.argv:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ia, isrc, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(r[ia]) for r in data); tots = max(1, sum(int(r[ist]) for r in data))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
        o = m.group(2).split('.')[0] if m else '?'
        op[o] += int(r[ia]); ops[o] += int(r[ist])
    print('  total warp instructions', tot)
    for o, c in op.most_common(16):
        print(f'  {o:10s} {c:12d} {c / tot * 100:5.1f}%   stall samples {ops[o] / tots * 100:5.1f}%')
