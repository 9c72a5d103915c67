#!/usr/bin/env python
"""Top stall-sample SASS lines of an `ncu --page source --csv` export, with the dominant stall reason per line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print(f"{rows[0][1][:80]}  total samples {tot}, SASS lines {len(body)}")
agg = {}
for r in body:
    for i in stall:
        if i < len(r) and r[i].isdigit(): agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("  by reason:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in enumerate(sorted(body, key=lambda r: -int(r[si]))[:top]):
    why = max(stall, key=lambda i: int(r[i]) if i < len(r) and r[i].isdigit() else 0)
    print(f"  {int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}%  exec {r[ii]:>10s}  {hdr[why][6:]:12s} {r[1].strip()[:90]}")
