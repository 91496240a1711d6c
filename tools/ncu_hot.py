#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ci["# Samples"]].isdigit()]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, "instructions", len(body))
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
top = sorted(enumerate(body), key=lambda ir: -int(ir[1][ci["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i, r in sorted(top):
    s = int(r[ci["# Samples"]] or 0)
    why = sorted(((int(r[ci[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  {r[ci['Source']].strip():60s} {why}")
