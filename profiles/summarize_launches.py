#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
python profiles/summarize_launches.py profiles/r01_launches_10m.csv > profiles/r01_launches_10m_summary.json"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= iV:
        continue
    v = float(r[iV].replace(",", ""))
    ms = v / 1e6 if r[iU] == "ns" else v / 1e3 if r[iU] == "us" else v
    name = re.sub(r"\(.*", "", r[iN]).replace("tess::<", "").replace("(bool)", "")
    tot[name] += ms
    cnt[name] += 1
s = sum(tot.values())
print(json.dumps([{"kernel": k, "launches": cnt[k], "total_ms": round(v, 4), "share": round(v / s, 5)} for k, v in tot.most_common()], indent=1))
