"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces them bit for bit (pins compiler / flags / refactors).
GPU (-m gpu): the CUDA path reproduces them through the C ABI with nothing under oracle/
involved -- bit for bit wherever no banana (atan2) or DEWPOINT (log) is in play, within the
north-star tolerance (1e-5 relative / 1e-3 absolute) otherwise."""
from pathlib import Path

import numpy as np
import pytest

from obsgrid_b200 import synth

G = Path(__file__).resolve().parent / "golden"
RTOL, ATOL = 1e-5, 1e-3
CASE_ARRAYS = ("pressure", "t", "u", "v", "rh", "slp", "xob", "yob", "lon", "elev", "ob_u", "ob_v",
               "ob_t", "ob_rh", "ob_slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")


def load(name):
    return np.load(G / name, allow_pickle=True)


def golden_case(z):
    a = z["args"].item()
    case = synth.small_case(**a)
    for n in CASE_ARRAYS:
        getattr(case, n)[...] = z["in_" + n]
    return case


class Impl:
    """The oracle and the GPU context expose the same method names; this picks one."""

    def __init__(self, kind, oracle=None, ctx=None):
        self.kind, self.o, self.c = kind, oracle, ctx

    def __getattr__(self, name):
        return getattr(self.o if self.kind == "oracle" else self.c, name)

    def proc_qc(self, case):
        self.o.qc_time(case) if self.kind == "oracle" else self.c.proc_qc(case)

    def proc_oa(self, case):
        self.o.oa_time(case) if self.kind == "oracle" else self.c.proc_oa(case)


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def impl(request):
    if request.param == "oracle":
        return Impl("oracle", oracle=request.getfixturevalue("oracle"))
    return Impl("gpu", ctx=request.getfixturevalue("ctx"))


def test_cressman_circle(impl):
    z = load("cressman_circle.npz")
    args = (z["d"], z["x"], z["y"], z["g"], int(z["var"]), int(z["R"]), float(z["dxd"]))
    np.testing.assert_array_equal(impl.cressman(*args), z["out"])
    np.testing.assert_array_equal(impl.cressman(*args, use_first_guess=False), z["out_nofg"])


def test_cressman_ellipse_and_banana(impl):
    z = load("cressman_aniso.npz")
    R, dxd = int(z["R"]), float(z["dxd"])
    out = impl.cressman(z["d"], z["x"], z["y"], z["g"], 1, R, dxd, z["ub"], z["vb"], 300.0)
    np.testing.assert_array_equal(out, z["out_ellipse"])          # circles + ellipses: exact
    out = impl.cressman(z["d"], z["x"], z["y"], z["g"], 4, R, dxd, z["ub2"], z["vb2"], 850.0,
                        fg_value=z["fg"], scale_rh=True)
    assert int(z["shapes_banana"][2]) > 0
    if impl.kind == "oracle":
        np.testing.assert_array_equal(out, z["out_banana"])
    else:                                                          # atan2: device vs glibc
        np.testing.assert_allclose(out, z["out_banana"], rtol=RTOL, atol=ATOL)


def test_oa_slab(impl):
    z = load("oa_slab.npz")
    out, rc, ns = impl.oa_slab(z["g"], 3, float(z["pressure"]), z["ob"], z["x"], z["y"],
                               tuple(int(r) for r in z["radii"]), float(z["dxd"]))
    assert rc == int(z["rc"]) and ns == int(z["nscan"])
    np.testing.assert_array_equal(out, z["out"])


def test_qc_slab(impl):
    z = load("qc_slab.npz")
    n = z["ob"].size
    q1, aob, err, thr = impl.error_max(z["ob"], z["x"], z["y"], z["elev"], np.zeros(n, np.int32),
                                       z["g"], int(z["var"]), float(z["max_error"]),
                                       float(z["pressure"]), z["lt"])
    np.testing.assert_array_equal(q1, z["qc_after_error_max"])
    np.testing.assert_array_equal(err, z["err_max"])
    np.testing.assert_array_equal(thr, z["thr_max"])
    q2, aob, errb, dm, bnf = impl.buddy_check(z["ob"], z["x"], z["y"], q1, z["elev"], z["g"],
                                              int(z["var"]), float(z["pressure"]), float(z["dx_km"]),
                                              1.0, float(z["max_buddy"]))
    np.testing.assert_array_equal(q2, z["qc_after_buddy"])
    np.testing.assert_array_equal(errb, z["err_buddy"])
    np.testing.assert_array_equal(dm, z["difmax"])
    np.testing.assert_array_equal(bnf, z["buddy_num_final"])
    assert np.count_nonzero(q2 & (1 << 17)) > 0 and np.count_nonzero(q1 & (1 << 16)) > 0


def test_time_small(impl):
    z = load("time_small.npz")
    case = golden_case(z)
    impl.proc_qc(case)
    # every flag, DEWPOINT-derived RH flags included: the device's LOG is glibc's logf restated
    for n in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(case, n), z[n], err_msg=n)
    impl.proc_oa(case)
    for n in ("t", "slp"):
        np.testing.assert_array_equal(getattr(case, n), z[n], err_msg=n)
    for n in ("u", "v", "rh"):
        if impl.kind == "oracle":
            np.testing.assert_array_equal(getattr(case, n), z[n], err_msg=n)
        else:
            np.testing.assert_allclose(getattr(case, n), z[n], rtol=RTOL, atol=ATOL, err_msg=n)


def test_epilogue_fdda_and_temporal_interp(impl):
    """SURVEY 8f ranks 2-4 (tests/golden/epilogue.npz): bit for bit, except qv / psfc on the GPU
    (EXP, LOG, ** in FP64-then-round vs glibc: 1e-6 relative, written here)."""
    z = load("epilogue.npz")
    gu, gv = impl.wind_increment_to_c_grid(z["u"], z["uA"], z["uC"], z["v"], z["vA"], z["vC"])
    np.testing.assert_array_equal(gu, z["out_u"]); np.testing.assert_array_equal(gv, z["out_v"])
    np.testing.assert_array_equal(impl.pres_sfc_from_slp(z["pres"], z["slpc"], z["slpx"]), z["out_pres"])
    cf, c1, c2 = (int(q) for q in z["tinterp"])
    for cross, key in ((True, "out_tinterp_cross"), (False, "out_tinterp_full")):
        out = impl.temporal_interp(z["ta"], z["tb"], cf, c1, c2, cross)
        np.testing.assert_array_equal(out[1] if impl.kind == "oracle" else out, z[key])
    out = impl.mixing_ratio(z["rh"], z["t"], z["h"], z["terrain"], z["slp"], z["press"])
    rh1, qv, psfc = out[-3:]
    np.testing.assert_array_equal(rh1, z["out_rh"])
    if impl.kind == "oracle":
        np.testing.assert_array_equal(qv, z["out_qv"]); np.testing.assert_array_equal(psfc, z["out_psfc"])
    else:
        np.testing.assert_allclose(qv, z["out_qv"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(psfc, z["out_psfc"], rtol=1e-6, atol=0)
    case = synth.small_case(**z["fdda_args"].item())
    for n in CASE_ARRAYS:
        getattr(case, n)[...] = z["fdda_in_" + n]
    tb, od = (impl.o.qc_time_fdda if impl.kind == "oracle" else impl.c.proc_qc_fdda)(case, z["fdda_tb0"])
    np.testing.assert_array_equal(tb, z["fdda_tb"]); np.testing.assert_array_equal(od, z["fdda_od"])
    for n in ("qc_u", "qc_v", "qc_t", "qc_slp"):
        np.testing.assert_array_equal(getattr(case, n), z["fdda_" + n], err_msg=n)
    assert np.count_nonzero(case.qc_rh != z["fdda_qc_rh"]) <= (0 if impl.kind == "oracle" else 2)
