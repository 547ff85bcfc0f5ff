"""The surface-FDDA composition (obsgrid_b200/fdda.py) with the CPU oracle as its backend: the order of
the driver's loop (src/driver.F90:107-141, :306-341) -- no GPU involved; the GPU chain is held to this
composition in tests/test_gpu_parity.py::test_surface_fdda_chain."""
import numpy as np

from obsgrid_b200 import fdda
from tests.test_gpu_parity import _OracleProc


def test_fdda_times_follow_the_driver():
    # interval / intf4d = 6: the first time has no FDDA pass of its own here (fdda_loop_max = 2 repeats the
    # analysis time); every later time is preceded by 5 FDDA times, the 6th (== the 3-D time) is skipped
    t = fdda.fdda_times(3, 6)
    assert len(t) == 2 * 5
    assert [x[1] for x in t[:5]] == [2, 3, 4, 5, 6] and all(x[2:] == (1, 7) for x in t[:5])
    assert [x[1] for x in t[5:]] == [8, 9, 10, 11, 12] and all(x[2:] == (7, 13) for x in t[5:])
    assert fdda.fdda_times(1, 6) == []


def test_chain_with_the_oracle_backend(oracle):
    calls = []

    class P(_OracleProc):
        def proc_qc_fdda(self, case, tob):
            assert not tob.any()                     # driver.F90:140-141: cleared at every analysis time
            calls.append(("qc", case.kbu))
            return super().proc_qc_fdda(case, tob)

        def proc_oa(self, case):
            calls.append(("oa", case.kbu, int(case.oa_params.surface_only)))
            return super().proc_oa(case)

    small = dict(iew=40, jns=36, kbu=3, n_sfc=100, n_snd=20, radii=(5, 3), dx_m=60000.0)
    cases3d, hourly = fdda.synthetic_chain(None, 3, 3, **small)
    before = [h.t.copy() for h in hourly]
    guesses = []

    class B(fdda.HostBackend):
        def analyse_surface(self, hc, tob):
            guesses.append(hc.t[0].copy())
            return super().analyse_surface(hc, tob)

    fdda.run_chain(B(P(oracle)), cases3d, hourly, 3)
    # 3-D time 1; 3-D time 2 then its two FDDA times; 3-D time 3 then its two
    kinds = [c[0] + str(c[1]) for c in calls]
    assert kinds == ["qc3", "oa3"] + ["qc3", "oa3", "qc1", "oa1", "qc1", "oa1"] * 2
    assert all(c[2] == 1 for c in calls if c[0] == "oa" and c[1] == 1)      # FDDA times: surface only
    # an FDDA time's first guess is the interpolation of the two analysed 3-D times, then analysed
    for h, b in zip(hourly, before):
        assert np.any(h.t != b)
    # the first guess of the first FDDA time: a convex combination (weights 2/3, 1/3) of the two analysed
    # 3-D times around it (tinterp_module.F90:126-213), cross points only
    a, b = cases3d[0].t[0], cases3d[1].t[0]
    lo, hi = np.minimum(a, b)[:-1, :-1], np.maximum(a, b)[:-1, :-1]
    g = guesses[0][:-1, :-1]
    assert np.all(g >= lo - 1e-3) and np.all(g <= hi + 1e-3)
    want = (a.astype(np.float64) * 2 + b.astype(np.float64)) / 3
    np.testing.assert_allclose(g, want[:-1, :-1], rtol=1e-5, atol=1e-3)
