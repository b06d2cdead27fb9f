"""Per-source-line totals of an .ncu-rep (warp instructions executed, stall samples): python tools/ncu_lines.py rep [N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None
hdr = None
tot_e = tot_s = 0
acc = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if hdr is None or not r or r[0] == '': continue
    try: e = int(r[ie]); s = int(r[isamp])
    except Exception: continue
    key = (cur, int(r[0]))
    a = acc.setdefault(key, [0, 0, r[1][:110]])
    a[0] += e; a[1] += s
    tot_e += e; tot_s += s
print('total exec', tot_e, 'samples', tot_s)
for (f, ln), (e, s, src) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f'{f}:{ln:5d} exec {e:10d} {100*e/tot_e:5.1f}%  samp {s:6d} {100*s/max(tot_s,1):5.1f}%  {src}')
