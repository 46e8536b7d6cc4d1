#!/usr/bin/env python
"""Per-kernel totals and the in-order launch sequence from an ncu --csv launch list (gpu__time_duration.sum)."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(sys.argv[1])) if r and r[0] == "ID")
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = OrderedDict()
seq = []
for r in rows:
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    name = r[ik][:90]
    seq.append((name, v))
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += v
allms = sum(t[1] for t in tot.values())
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-92s n=%4d total_ms=%9.3f avg_ms=%8.4f share=%5.1f%%" % (name, n, ms, ms / n, 100 * ms / allms))
if len(sys.argv) > 2:
    print("# launch sequence (last %s)" % sys.argv[2])
    for name, v in seq[-int(sys.argv[2]):]:
        print("  %8.4f ms  %s" % (v, name))
