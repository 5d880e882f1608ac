#!/usr/bin/env python
"""Join tools/reorder_probe.py --profile rows with the ncu CSV of the same run (one captured launch per row, in order).

    python tools/reorder_merge_ncu.py probe.json ncu.csv out.json
"""
import csv
import json
import sys


def main():
    probe = json.load(open(sys.argv[1]))
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    ii, ni, vi, ui, ki = (hdr.index(x) for x in ("ID", "Metric Name", "Metric Value", "Metric Unit", "Kernel Name"))
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[ii]), {"kernel": r[ki].split("(")[0]})
        try:
            d[r[ni]] = (float(r[vi].replace(",", "")), r[ui])
        except ValueError:
            pass
    seq = [launches[k] for k in sorted(launches)]
    assert len(seq) == len(probe["rows"]), (len(seq), len(probe["rows"]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row, l in zip(probe["rows"], seq):
        row["ncu_kernel"] = l["kernel"]
        row["l2_sector_hit_rate_pct"] = round(l["lts__t_sector_hit_rate.pct"][0], 2)
        v, u = l["dram__bytes_read.sum"]
        row["dram_read_gb"] = round(v * scale[u] / 1e9, 3)
    json.dump(probe, open(sys.argv[3], "w"), indent=1)
    for r in probe["rows"]:
        print(r["order"], r["ef"], r["qps"], r["l2_sector_hit_rate_pct"], r["dram_read_gb"])


if __name__ == "__main__":
    main()
