"""Summarise an `ncu --csv` launch log (gpu__time_duration.sum per launch) into per-kernel totals."""
import csv
import collections
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    rows.append((name, ns, r["Kernel Name"]))
rows = rows[skip:]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for n, ns, _ in rows:
    tot[n] += ns
    cnt[n] += 1
total = sum(tot.values())
print("launches {}  total {:.1f} us".format(len(rows), total / 1e3))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("{:8.1f} us  {:5.1f}%  n={:4d}  avg {:7.2f} us  {}".format(t / 1e3, 100 * t / total, cnt[n], t / cnt[n] / 1e3, n[:90]))
