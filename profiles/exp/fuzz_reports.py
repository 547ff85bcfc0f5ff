#!/usr/bin/env python
"""CPU-only fuzz of the report list <-> table conversion against the as-written restatement
(oracle/og_oracle_reports.c): many seeds, both pressure tolerances.  Prints the first mismatch.

    python profiles/exp/fuzz_reports.py [n_seeds]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from obsgrid_b200 import reports as R, synth      # noqa: E402
from oracle import ogoracle                       # noqa: E402
from tests import test_reports as T               # noqa: E402


def main(n):
    bad = shared_cases = 0
    for seed in range(100, 100 + n):
        for p_diff in (1, 1500):
            kbu = int(np.random.default_rng(seed).integers(3, 12))
            case = synth.small_case(iew=60, jns=50, kbu=kbu, n_sfc=300, n_snd=120, radii=(5, 3), seed=seed)
            rl, date, time = R.synthetic_reports(case, seed=seed * 7 + 1)
            a, b = rl.copy(), rl.copy()
            if R.store_ob_defect_reports(rl, case.pressure):     # the reference's store_ob defect: not reproduced
                print("  (store_ob defect input, skipped) seed", seed)
                continue
            try:
                assert ogoracle.qc_time_reports(case, a, date, time, p_diff) == 0
                T._use_requests_of_proc_oa(ogoracle, case, a, date, time, p_diff)
                t = T._qc_through_table(case, b, date, time, p_diff, lambda c: ogoracle.qc_time(c))
                shared = R.shared_levels(t)
                shared_cases += 1 if shared else 0
                T._assert_lists_equal(a, b)
            except AssertionError as e:
                if shared:          # one measurement answering two levels: exact with the consistency pass left to the list
                    c = rl.copy()
                    T._qc_through_table(case, c, date, time, p_diff, lambda q: ogoracle.qc_time(q, run_consistency=False))
                    T._assert_lists_equal(a, c)
                    print("  (per-row consistency differs, list consistency exact) seed", seed, "p_diff", p_diff, "shared entries", shared)
                    continue
                bad += 1
                print("MISMATCH seed", seed, "p_diff", p_diff, "kbu", kbu, str(e)[:400])
                if bad > 3:
                    return bad
    print("seeds", n, "x 2 tolerances: mismatches", bad, "| cases with shared levels:", shared_cases)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 20) else 0)
