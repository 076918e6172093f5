"""Mnemonic histogram + excerpt of one kernel's SASS from libqcknot.so (cuobjdump -sass): evidence for profiles/."""
import collections, re, subprocess, sys
lib = sys.argv[1]
pat = sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if "Function :" in l and re.search(pat, l))
end = next((i for i in range(start + 1, len(out)) if "Function :" in out[i]), len(out))
body = out[start:end]
ops = collections.Counter()
for l in body:
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        parts = m.group(2).split(".")
        keep2 = parts[0] in ("UBLKCP", "LDS", "STS", "STG", "LDG", "DMMA", "BAR", "ST", "LD") and len(parts) > 1
        ops[parts[0] + ("." + ".".join(parts[1:3]) if keep2 else "")] += 1
print(body[0].strip())
print("static SASS instructions:", sum(ops.values()))
for k, v in ops.most_common(30):
    print(f"  {k:16s} {v}")
print("-- first UBLKCP sites:")
n = 0
for i, l in enumerate(body):
    if "UBLKCP" in l:
        print("   ", body[i].strip()[:110])
        n += 1
        if n >= 4:
            break
