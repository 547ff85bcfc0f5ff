#!/usr/bin/env python
"""Per-kernel totals of the LAST step in an ncu `--metrics gpu__time_duration.sum --csv` launch list.

    python profiles/launch_summary.py gpurun_out/launches_cfg3.csv [steps_in_file]
"""
import csv, sys, collections, re

def main(path, steps=2):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
    rows = [r for r in rows if "at::native" not in r[0] and "elementwise" not in r[0]]
    n = len(rows) // steps
    last = rows[-n:]
    tot = collections.OrderedDict()
    for k, us in last:
        k = re.sub(r"\(.*", "", k)
        a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    total = sum(v[1] for v in tot.values())
    for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} x{c:4d} {us:10.1f} us {100*us/total:5.1f}%")
    print(f"total {total:.1f} us in {n} launches")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
