#!/usr/bin/env python
"""Hottest SASS instructions of an `ncu --page source --csv --print-source sass` export: stall samples per instruction with the dominant
stall reasons.  usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_hot.py x.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for k, r in enumerate(rows[2:]):
    try:
        n = int(r[idx["# Samples"]])
    except (ValueError, IndexError):
        continue
    data.append((n, k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for n, k, r in data:
    for s in stalls:
        try:
            agg[s] = agg.get(s, 0) + int(r[idx[s]])
        except ValueError:
            pass
print("by reason:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, k, r in sorted(data, reverse=True)[:top]:
    rs = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print("%5d %5.1f%%  #%4d  %-70s  %s  exec=%s" % (n, 100.0 * n / max(tot, 1), k, r[idx["Source"]][:70], " ".join("%s:%d" % (b, a) for a, b in rs if a), r[idx["Instructions Executed"]]))
