"""Rank SASS instructions of an ncu report by warp-stall samples (ncu -i X.ncu-rep --page source --csv > X.csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci['# Samples']]), n, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for s, n, r in sorted(data, reverse=True)[:top]:
    why = sorted(((int(r[ci[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{s:6d} {100*s/tot:5.1f}% #{n:5d} {r[1].strip()[:70]:70s} {why}")
# opcode-class summary
cls = {}
for s, n, r in data:
    op = r[1].strip().split()[0] if r[1].strip() else '?'
    if op.startswith('@'):
        op = r[1].strip().split()[1]
    op = op.split('.')[0]
    cls[op] = cls.get(op, 0) + s
print(sorted(cls.items(), key=lambda kv: -kv[1])[:12])
