"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) per kernel: python tools/summarize_launches.py file.csv [steps]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("lavt::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else v * 1e3 if r[iu] == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/steps:10.1f} us {100*v[1]/tot:5.1f}%  x{v[0]/steps:6.1f}  {k[:100]}")
print(f"{tot/steps:10.1f} us total per step")
