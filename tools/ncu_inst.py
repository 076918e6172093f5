"""Executed warp-instruction histogram by opcode and by code region (ncu --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
tot = 0; cls = {}; data = []
for n, r in enumerate(rows[2:]):
    try: ex = int(r[ci['Instructions Executed']])
    except Exception: continue
    txt = r[1].strip(); parts = txt.split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    op = op.split('.')[0]
    cls[op] = cls.get(op, 0) + ex; tot += ex; data.append((n, ex, txt))
print("total warp instr", tot)
for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:25]: print(f"  {k:10s} {v:12d} {100*v/tot:5.1f}%")
# region profile: bucket of 200 instructions
B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
for b in range(0, len(data), B):
    chunk = data[b:b+B]
    s = sum(e for _, e, _ in chunk)
    dfma = sum(e for _, e, t in chunk if 'DFMA' in t or 'DMUL' in t or 'DADD' in t)
    stg = sum(e for _, e, t in chunk if 'STG' in t)
    lds = sum(e for _, e, t in chunk if 'LDS' in t)
    bar = sum(1 for _, e, t in chunk if 'BAR.SYNC' in t)
    print(f"  #{b:5d}-{b+B:5d}: {s:11d} ({100*s/tot:4.1f}%) f64={100*dfma/max(s,1):4.0f}% lds={100*lds/max(s,1):3.0f}% stg={100*stg/max(s,1):3.0f}% bars={bar}")
