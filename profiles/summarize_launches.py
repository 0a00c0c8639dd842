"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and launch count per kernel name.
    python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"(anonymous namespace)::", "", name.replace("(anonymous namespace)::", ""))
    try:
        t = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[name] += t
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k] / 1e3:10.1f} {tot[k] / cnt[k] / 1e3:8.2f} {100 * tot[k] / total:5.1f}%")
print(f"{'TOTAL':70s} {sum(cnt.values()):8d} {total / 1e3:10.1f}")
