"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":   # (the same log may carry DRAM byte counters)
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0 if u in ("ms", "msecond") else v)
        rows.append((r[ki], us))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 1e30])
    for k, us in rows:
        name = re.sub(r"\(.*", "", k).replace("void ", "").replace("valor::", "")
        agg[name][0] += 1
        agg[name][1] += us
        agg[name][2] = max(agg[name][2], us)
        agg[name][3] = min(agg[name][3], us)
    total = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}  total {total/1000:.2f} ms")
    for name, (n, us, mx, mn) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{us/1000:9.3f} ms {100*us/total:5.1f}%  n={n:5d}  avg {us/n:9.1f} us  min {mn:8.1f} max {mx:9.1f}  {name[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
