"""Summarise an `ncu --page source --csv` dump: instruction mix and stall reasons (development tool)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
data = []
for r in rows[hdr + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break                      # next captured launch: keep the first one only
    if len(r) == len(H):
        data.append(r)
si, ii, ni = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_")]
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = 0
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
    op = m.group(2).split(".")[0] if m else "?"
    n = int(float(r[ii] or 0)); tot += n
    ops[op] += n; samp[op] += int(float(r[ni] or 0))
    for c in stall_cols:
        stalls[H[c]] += int(float(r[c] or 0))
print("total warp instructions:", tot)
for op, n in ops.most_common(22):
    print(f"  {op:10s} {n:14d} {100*n/tot:5.1f}%   samples {samp[op]}")
ts = sum(stalls.values())
print("stall samples:", {k: f"{100*v/ts:.1f}%" for k, v in stalls.most_common(8)})
if len(sys.argv) > 2:
    top = sorted(data, key=lambda r: -int(float(r[ni] or 0)))[:int(sys.argv[2])]
    for r in top:
        print(f"  {r[ni]:>6s} {r[ii]:>12s}  {r[si][:90]}")
