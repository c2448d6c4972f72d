"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel (for profiles/)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
    name = r[kn][:96]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("launches %d, total device time %.1f us (serialised, cold-cache: compare SHARES)" % (sum(cnt.values()), T))
ours = sum(v for n, v in tot.items() if "dcfp::" in n)
print("dcfp kernels: %.1f us = %.2f %% of the region" % (ours, 100 * ours / T))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, v in tot.most_common(top):
    print("%9.1f us %5.1f%% x%4d  %s" % (v, 100 * v / T, cnt[n], n))
print("-- dcfp kernels")
for n, v in tot.most_common():
    if "dcfp::" in n:
        print("%9.1f us %5.2f%% x%4d  %s" % (v, 100 * v / T, cnt[n], n))
