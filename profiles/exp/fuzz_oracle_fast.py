#!/usr/bin/env python
"""CPU-only fuzz of the oracle's fast paths (binned buddy check, row-banded Cressman, ogo_set_fast_paths)
against the as-written loops over random geometries, densities, variables, pressure classes and thread
counts -- the fast paths are what checks the GPU on the 200 k - 2 M report slabs.

    python profiles/exp/fuzz_oracle_fast.py [n_cases]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import ogoracle as O                               # noqa: E402
from tests.test_oracle_fast_paths import _buddy_case           # noqa: E402


def main(n_cases):
    bad = 0
    for case in range(n_cases):
        rng = np.random.default_rng(1000 + case)
        iew, jns = int(rng.integers(20, 260)), int(rng.integers(20, 260))
        n = int(rng.choice([0, 1, 2, 3, 50, 400, 2500, 6000]))
        var = int(rng.choice([1, 2, 3, 4, 5, 6, 7]))
        pressure = float(rng.choice([1001.0, 1000.0, 850.0, 700.0, 500.0, 400.0, 300.0, 201.0, 200.0, 100.0]))
        dx = float(rng.choice([3.0, 9.0, 12.0, 30.0, 60.0, 150.0]))
        g, x, y, ob, elev = _buddy_case(rng, iew, jns, n, bool(rng.integers(0, 2)), float(rng.choice([3.0, 8.0, 300.0])))
        if var == 7:
            ob = (ob - 45).astype(np.float32)
        q0 = (rng.integers(0, 2, n) * rng.choice([4, 1 << 14, 1 << 16])).astype(np.int32)
        weight, mb = float(rng.choice([1.0, 0.7])), float(rng.choice([4.0, 8.0, 500.0]))
        O.set_fast_paths(False, 1)
        a = O.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, weight, mb)
        O.set_fast_paths(True, int(rng.integers(1, 9)))
        b = O.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, weight, mb)
        for nm, u, v in zip(("qc", "aob", "err", "difmax", "buddy_num_final"), a, b):
            if not np.array_equal(u, v):
                bad += 1
                print("BUDDY MISMATCH case", case, nm, dict(iew=iew, jns=jns, n=n, var=var, p=pressure, dx=dx))
                break
        # Cressman, one scan
        if n > 0:
            R = int(rng.choice([1, 2, 4, 7, 12]))
            ub = (30 * np.sin(np.arange(jns)[:, None] / 20.0) + 5 + 0 * g).astype(np.float32)
            vb = (25 * np.cos(np.arange(iew)[None, :] / 25.0) + 0 * g).astype(np.float32)
            d = rng.normal(0, 5, n).astype(np.float32)
            cvar = int(rng.choice([1, 2, 3, 4, 5]))
            O.set_fast_paths(False, 1)
            ca = O.cressman(d, x, y, g, cvar, R, dx, u_banana=ub, v_banana=vb, pressure=pressure)
            O.set_fast_paths(True, int(rng.integers(1, 9)))
            cb = O.cressman(d, x, y, g, cvar, R, dx, u_banana=ub, v_banana=vb, pressure=pressure)
            if not np.array_equal(ca, cb):
                bad += 1
                print("CRESSMAN MISMATCH case", case, dict(iew=iew, jns=jns, n=n, var=cvar, R=R, dx=dx))
    O.set_fast_paths(False, 1)
    print("cases", n_cases, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 40) else 0)
