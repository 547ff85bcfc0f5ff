#!/usr/bin/env python
"""CPU-only fuzz of the C oracle (oracle/og_oracle.c) against the independent numpy restatement written
from the Fortran source (tests/ref_numpy.py): random grids, winds, radii, variables, pressure classes.
Cressman fields must be identical wherever no banana contributes (numpy's arctan2 is not glibc's);
buddy_check and error_max must be identical always.

    python profiles/exp/fuzz_oracle_vs_numpy.py [n_cases]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import ogoracle as O                                    # noqa: E402
from tests import ref_numpy as R                                    # noqa: E402
from tests.test_oracle_vs_numpy import wavy, winds, NAMES           # noqa: E402


def main(n_cases):
    bad = 0
    for case in range(n_cases):
        rng = np.random.default_rng(5000 + case)
        iew, jns = int(rng.integers(12, 70)), int(rng.integers(12, 60))
        n = int(rng.choice([1, 2, 5, 40, 120]))
        x = rng.uniform(1.0, iew - 0.2, n).astype(np.float32); y = rng.uniform(1.0, jns - 0.2, n).astype(np.float32)
        # ---- Cressman --------------------------------------------------------------------------
        var = int(rng.choice([1, 2, 3, 4, 5]))
        Rr = int(rng.choice([1, 2, 3, 5, 8, 12]))
        pressure = float(rng.choice([1001.0, 850.0, 700.0, 500.0, 300.0, 150.0]))
        g = wavy(iew, jns, 20.0, 50.0)
        ub, vb = winds(iew, jns, jet=float(rng.choice([0.0, 8.0, 40.0, 90.0])), curl=bool(rng.integers(0, 2)))
        d = rng.normal(0, 5, n).astype(np.float32)
        fg = (np.abs(rng.normal(50, 20, n)) + 1).astype(np.float32)
        ufg, srh = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        dx = float(rng.choice([3.0, 12.0, 30.0, 90.0]))
        O.reset_stats()
        a = O.cressman(d, x, y, g, var, Rr, dx, ub, vb, pressure, fg_value=fg, scale_rh=srh, use_first_guess=ufg)
        st = O.get_stats()
        b, shapes = R.cressman(d, x, y, g, NAMES[var], Rr, dx, ub, vb, pressure, fg_value=fg, scale_rh=srh,
                               use_first_guess=ufg)
        ok = shapes == [st["n_circle"], st["n_ellipse"], st["n_banana"]]
        ok = ok and (np.array_equal(a, b) if shapes[2] == 0 else np.allclose(a, b, rtol=1e-5, atol=1e-3))
        if not ok:
            bad += 1
            print("CRESSMAN MISMATCH case", case, dict(iew=iew, jns=jns, n=n, var=var, R=Rr, p=pressure, dx=dx), shapes,
                  [st["n_circle"], st["n_ellipse"], st["n_banana"]])
        # ---- buddy_check / error_max ---------------------------------------------------------------
        qv = int(rng.choice([1, 2, 3, 4, 5, 6, 7]))
        x2 = np.clip(x, 2.0, iew - 1.0).astype(np.float32); y2 = np.clip(y, 2.0, jns - 1.0).astype(np.float32)
        base = {1: 10.0, 2: 10.0, 3: 270.0, 4: 60.0, 5: 101000.0, 6: 101000.0, 7: 240.0}[qv]
        spread = {5: 400.0, 6: 400.0, 7: 25.0}.get(qv, 6.0)
        gq = wavy(iew, jns, spread / 3, base)
        ob = (base + rng.normal(0, spread, n)).astype(np.float32)
        ob[rng.random(n) < 0.1] += np.float32(4 * spread)
        elev = rng.uniform(-3000, 3500, n).astype(np.float32)
        lt = (rng.integers(0, 24, n) * 100).astype(np.int32)
        q0 = (rng.integers(0, 2, n) * (1 << 14) + rng.integers(0, 2, n) * (1 << 16) + rng.integers(0, 2, n) * 4).astype(np.int32)
        maxd = {5: 300.0, 6: 300.0, 7: 12.0}.get(qv, 5.0)
        ea = O.error_max(ob, x2, y2, elev, q0, gq, qv, maxd, pressure, lt)
        eb = R.error_max(ob, x2, y2, elev, q0, gq, NAMES[qv], maxd, pressure, lt)
        if not all(np.array_equal(u, v) for u, v in zip(ea, eb)):
            bad += 1
            print("ERROR_MAX MISMATCH case", case, dict(var=qv, p=pressure))
        w = float(rng.choice([1.0, 0.8]))
        ba = O.buddy_check(ob, x2, y2, q0, elev, gq, qv, pressure, dx, w, maxd)
        bb = R.buddy_check(ob, x2, y2, q0, elev, gq, NAMES[qv], pressure, dx, w, maxd)
        if not all(np.array_equal(u, v) for u, v in zip(ba, bb)):
            bad += 1
            print("BUDDY MISMATCH case", case, dict(var=qv, p=pressure, dx=dx, n=n))
    print("cases", n_cases, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 40) else 0)
