"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name and grid."""
import collections
import csv
import io
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.defaultdict(lambda: [0, 0.0])
fam = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:64]
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[(n, r["Grid Size"])][0] += 1
    agg[(n, r["Grid Size"])][1] += t
    f = re.sub(r"<.*", "", n)
    fam[f][0] += 1
    fam[f][1] += t
tot = sum(v[1] for v in agg.values())
print(f"launches {sum(v[0] for v in agg.values())}, total {tot / 1e3:.2f} ms")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:44s} {v[0]:5d} {v[1] / 1e3:8.3f} ms {100 * v[1] / tot:5.1f}%")
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:66s} {k[1]:16s} {v[0]:4d} {v[1]:9.1f} us {v[1] / v[0]:8.1f}")
