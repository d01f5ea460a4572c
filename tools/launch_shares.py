"""Kernel shares from an ncu launch list (--metrics gpu__time_duration.sum --csv):  python tools/launch_shares.py launches.csv [skip]"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1 + skip:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("us", "usecond") else (v / 1e6 if r[ui] in ("ns", "nsecond") else v)
    name = r[ki].split("(")[0].replace("void ", "").strip()
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':44s} launches   total ms   ms/launch   share")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:44s} {cnt[k]:8d} {v:10.3f} {v / cnt[k]:10.4f} {100 * v / s:6.1f} %")
