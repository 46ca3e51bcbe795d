#!/usr/bin/env python
"""Write / update profiles/ncu_traffic.json from `ncu -i X.ncu-rep --page raw --csv` dumps.

    python profiles/ncu_traffic.py --grid 8192x8192 profiles/r02_raw_rbgs_stream.csv [more.csv ...]

For every launch in the dumps: dram__bytes_read.sum + dram__bytes_write.sum -> kernels[<bench.py kernel name>].  bench.py
reads this file for `roofline.traffic` (no constants in bench.py).  A PCG iteration is the sum of its two kernels.
"""
import argparse
import csv
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    """'void k_rhs_march<1>(Geom, ...)' -> 'k_rhs_march<1>'; 'k_rbgs_stream<1, 0>' -> 'k_rbgs_stream'."""
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", name)
    base, targs = m.group(1), m.group(2) or ""
    if base in ("k_rhs_march", "k_rhs_march2", "k_divergence", "k_rhs_fused2"):
        return base + targs.replace(" ", "")
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", required=True)
    ap.add_argument("csv", nargs="+")
    a = ap.parse_args()
    data = {"grid": a.grid, "kernels": {}}
    if os.path.exists(OUT):
        old = json.load(open(OUT))
        if old.get("grid") == a.grid:
            data = old
    for path in a.csv:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
            data["kernels"][short(r[ik])] = {"dram_bytes": b, "source": "profiles/" + os.path.basename(path)}
    k = data["kernels"]
    if "k_pcg_phase1_march" in k and "k_pcg_phase2_march" in k:
        k["k_pcg_phase1_march + k_pcg_phase2_march (one iteration)"] = {
            "dram_bytes": k["k_pcg_phase1_march"]["dram_bytes"] + k["k_pcg_phase2_march"]["dram_bytes"],
            "source": k["k_pcg_phase1_march"]["source"] + " + " + k["k_pcg_phase2_march"]["source"]}
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
