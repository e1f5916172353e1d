#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name: launches, total / mean us, share.

    python tools/ncu_summary.py launches.csv [skip_first_n_launches]
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        rows.append((r[ki], us))
    rows = rows[skip:]
    agg = OrderedDict()
    for k, us in rows:
        k = re.sub(r"\(.*\)$", "", k)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot:.1f} us total (serialised, cold cache: shares matter, not absolutes)")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={n:4d}  mean {us / n:8.1f} us  {k[:110]}")


if __name__ == "__main__":
    main()
