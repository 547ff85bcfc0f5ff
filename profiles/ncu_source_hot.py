#!/usr/bin/env python
"""Top source lines of an `ncu --page source --print-source cuda,sass --csv` export:
share of executed warp instructions and of stall samples per CUDA line."""
import csv
import sys


def f(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and r[2] == "-":      # a CUDA line aggregate row
            lines.append((cur_file, r))
    ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    sidx = [hdr.index(h) for h in names]
    tot_i = sum(f(r[ci]) for _, r in lines); tot_s = sum(f(r[cs]) for _, r in lines)
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    agg = [0.0] * len(names)
    for _, r in lines:
        for k, i in enumerate(sidx):
            agg[k] += f(r[i])
    print("stall mix:", ", ".join(f"{n[6:]} {100 * a / max(sum(agg), 1):.1f}%" for n, a in
                                  sorted(zip(names, agg), key=lambda t: -t[1])[:8]))
    for fn, r in sorted(lines, key=lambda t: -f(t[1][cs]))[:top]:
        st = sorted(((f(r[i]), n[6:]) for n, i in zip(names, sidx)), reverse=True)[:2]
        print(f"{fn:14s} L{r[0]:>4} inst {100 * f(r[ci]) / tot_i:5.1f}% samp {100 * f(r[cs]) / tot_s:5.1f}% "
              f"[{st[0][1]},{st[1][1]}] {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
