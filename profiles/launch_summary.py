#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, average, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik])
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {sum(cnt.values())} launches captured, {total / 1e3:.3f} ms of kernel time")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k:44s} n={cnt[k]:4d} total={tot[k] / 1e3:9.3f} ms avg={tot[k] / cnt[k]:9.1f} us share={100 * tot[k] / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
