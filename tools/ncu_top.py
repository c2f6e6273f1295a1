"""Top stalled SASS instructions per kernel of an `ncu --page source --csv` dump.  Usage: python tools/ncu_top.py file.csv [kernel-index] [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
kern, cur = [], None
for r in rows:
    if len(r) >= 1 and r[0] == "Kernel Name":
        cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if cur is not None:
        cur['rows'].append(r)
print(len(kern), "kernels:", [k['name'][:40] for k in kern])
k = kern[which]
hdr, data = k['rows'][0], k['rows'][1:]
si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
tot = sum(int(r[si]) for r in data if r[si].isdigit())
print(k['name'][:80], 'total samples', tot, 'instrs', len(data))
top = sorted([(int(r[si]), i) for i, r in enumerate(data) if r[si].isdigit()], reverse=True)[:n]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
for s, i in top:
    r = data[i]
    st = sorted([(int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0], reverse=True)[:3]
    print(f"{s:6d} {100.0*s/tot:5.1f}% #{i:5d} ie={r[ie]:>7s} {r[src].strip()[:64]:64s} {st}")
