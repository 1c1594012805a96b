#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, share, launches)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        name = re.sub(r"^void ", "", d["Kernel Name"])
        name = re.sub(r"muse::<unnamed>::", "", name)
        name = re.sub(r"\(.*", "", name)[:90]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u.startswith("us") else v)
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.3f} ms over {len(data)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
