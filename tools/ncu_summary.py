"""Summarise an .ncu-rep (raw metrics + SASS opcode mix + stall reasons) - used to write profiles/*.md."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.avg.per_cycle_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f'{w} [{units[i]}]:', [r[i][:90] for r in rows[2:]])
fp = [h for h in hdr if 'fp64' in h]
for w in fp:
    i = hdr.index(w)
    print(f'{w} [{units[i]}]:', [r[i] for r in rows[2:]])
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
h = rows[1]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
ops, samp = collections.Counter(), collections.Counter()
tot = ts = n = 0
for r in rows[2:]:
    try:
        e, s = int(r[ie]), int(r[isamp])
    except Exception:
        continue
    op = r[ia].strip().split()
    if not op:
        continue
    o = (op[1] if op[0].startswith('@') else op[0]).split('.')[0]
    ops[o] += e; samp[o] += s; tot += e; ts += s; n += 1
print('static SASS instructions', n, 'executed warp-instructions', tot, 'stall samples', ts)
for o, c in ops.most_common(22):
    print(f'  {o:10s} exec {c:10d} {100 * c / tot:5.1f}%   samples {samp[o]:6d} {100 * samp[o] / max(ts, 1):5.1f}%')
st = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
totst = collections.Counter()
for r in rows[2:]:
    for i in st:
        try:
            totst[h[i]] += int(r[i])
        except Exception:
            pass
ssum = sum(totst.values())
print('stall reasons:', ', '.join(f'{k[6:]} {100 * v / max(ssum, 1):.1f}%' for k, v in totst.most_common(10)))
