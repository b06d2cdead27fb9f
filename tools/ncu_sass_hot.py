"""Hottest SASS instructions (stall samples) of an .ncu-rep with their source line: python tools/ncu_sass_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; line = None; items = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); st = [i for i, c in enumerate(hdr) if c.startswith('stall_')]; continue
    if hdr is None or not r: continue
    if r[0] != '': line = (cur, r[0], r[1][:70]); continue
    try: e = int(r[ie]); s = int(r[isamp])
    except Exception: continue
    top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in st if r[i] not in ('', '0')), reverse=True)[:2]
    items.append((s, e, r[3].strip()[:60], line, top))
tot = sum(i[0] for i in items)
for s, e, sass, line, top in sorted(items, key=lambda x: -x[0])[:N]:
    print(f'{s:5d} {100*s/tot:4.1f}% exec {e:8d}  {sass:60s} {line[0]}:{line[1]} {top}')
