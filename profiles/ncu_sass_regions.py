#!/usr/bin/env python
"""Segment an `ncu --page source --print-source sass --csv` export into runs of SASS instructions
with (nearly) equal execution counts -- loop bodies show up as plateaus -- and print each run's
share of the executed warp instructions with its opcode mix."""
import csv
import sys
from collections import Counter


def f(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, min_share=0.004):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ci, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ins = [r for r in rows if r and r[0].startswith("0x")]
    tot = sum(f(r[ci]) for r in ins)
    tots = sum(f(r[ss]) for r in ins)
    print(f"{len(ins)} SASS instructions, {tot:.3e} executed, {tots:.0f} samples")
    seg, cur, prev = [], [], None
    for r in ins:
        c = f(r[ci])
        if prev is not None and abs(c - prev) > 0.03 * max(c, prev, 1):
            seg.append(cur); cur = []
        cur.append(r); prev = c
    seg.append(cur)
    base = int(ins[0][0], 16)
    for s in seg:
        c = sum(f(x[ci]) for x in s)
        if c / tot < min_share:
            continue
        ops = Counter()
        for x in s:
            t = x[si].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
        sm = sum(f(x[ss]) for x in s)
        print(f"+{int(s[0][0], 16) - base:05x}..+{int(s[-1][0], 16) - base:05x} n={len(s):4d} "
              f"cnt={f(s[0][ci]):.3e} inst {100 * c / tot:5.1f}% samp {100 * sm / tots:5.1f}%  "
              f"{dict(ops.most_common(7))}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.004)
