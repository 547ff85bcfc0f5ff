"""The C oracle against a second, independent restatement (tests/ref_numpy.py, written from the
Fortran source).  With no Fortran toolchain the reference itself cannot run here; two
independent transcriptions agreeing bit for bit is the strongest pin available."""
import numpy as np
import pytest

from tests import ref_numpy as R

NAMES = {1: "UU", 2: "VV", 3: "TT", 4: "RH", 5: "PMSL", 6: "PMSLPSFC", 7: "DEWPOINT"}


def wavy(iew, jns, a=10.0, b=280.0):
    i = np.arange(iew)[None, :]; j = np.arange(jns)[:, None]
    return (b + a * np.sin(i / 7.0) * np.cos(j / 5.0)).astype(np.float32)


def winds(iew, jns, jet=40.0, curl=True):
    i = np.arange(1, iew + 1, dtype=np.float64)[None, :]
    j = np.arange(1, jns + 1, dtype=np.float64)[:, None]
    yc = jns / 2 + 0.15 * jns * np.sin(2 * np.pi * i / iew * 1.5)
    core = np.exp(-((j - yc) / (0.12 * jns)) ** 2)
    u = 3.0 + jet * core
    v = jet * core * (0.15 * jns * 2 * np.pi * 1.5 / iew) * np.cos(2 * np.pi * i / iew * 1.5) if curl else 0 * u
    return u.astype(np.float32), (v + 0 * u).astype(np.float32)


@pytest.mark.parametrize("var,R_,ufg", [(3, 5, True), (5, 3, False), (3, 9, True)])
def test_cressman_isotropic(oracle, var, R_, ufg):
    rng = np.random.default_rng(var * 10 + R_)
    iew, jns, n = 44, 37, 120
    g = wavy(iew, jns)
    x = rng.uniform(1.0, iew - 0.2, n).astype(np.float32); y = rng.uniform(1.0, jns - 0.2, n).astype(np.float32)
    d = rng.normal(0, 3, n).astype(np.float32)
    a = oracle.cressman(d, x, y, g, var, R_, 12.0, use_first_guess=ufg)
    b, _ = R.cressman(d, x, y, g, NAMES[var], R_, 12.0, use_first_guess=ufg)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("var,pressure,curl", [(1, 300.0, False), (2, 850.0, True), (4, 700.0, True), (4, 1001.0, True)])
def test_cressman_anisotropic(oracle, var, pressure, curl):
    rng = np.random.default_rng(int(pressure) + var)
    iew, jns, n, R_ = 64, 52, 90, 4
    g = wavy(iew, jns, 20.0, 50.0)
    ub, vb = winds(iew, jns, curl=curl)
    x = rng.uniform(1.0, iew - 0.2, n).astype(np.float32); y = rng.uniform(1.0, jns - 0.2, n).astype(np.float32)
    d = rng.normal(0, 5, n).astype(np.float32)
    fg = (np.abs(rng.normal(50, 20, n)) + 1).astype(np.float32)
    oracle.reset_stats()
    a = oracle.cressman(d, x, y, g, var, R_, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    st = oracle.get_stats()
    b, shapes = R.cressman(d, x, y, g, NAMES[var], R_, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    assert shapes == [st["n_circle"], st["n_ellipse"], st["n_banana"]]
    assert shapes[1] + shapes[2] > 0
    if shapes[2] == 0:
        np.testing.assert_array_equal(a, b)
    else:   # numpy's float32 arctan2 need not be glibc's atan2f to the last bit
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-3)
        assert np.mean(a != b) < 0.2 and np.max(np.abs(a - b)) < 1e-4


@pytest.mark.parametrize("var,pressure,n,dx", [(1, 850.0, 150, 30.0), (3, 300.0, 200, 60.0), (6, 1001.0, 120, 30.0),
                                               (7, 500.0, 100, 90.0), (5, 1001.0, 60, 150.0)])
def test_buddy_check(oracle, var, pressure, n, dx):
    rng = np.random.default_rng(var + n)
    iew, jns = 70, 60
    g = wavy(iew, jns, 6.0, 260.0 if var in (3, 7) else 10.0)
    x = rng.uniform(2.0, iew - 1.0, n).astype(np.float32); y = rng.uniform(2.0, jns - 1.0, n).astype(np.float32)
    x[1], y[1] = x[0], y[0]
    ob = (g.mean() + rng.normal(0, 5, n)).astype(np.float32)
    ob[rng.random(n) < 0.08] += 40.0
    elev = rng.uniform(-3000, 3000, n).astype(np.float32)
    q0 = (rng.integers(0, 2, n) * (1 << 16)).astype(np.int32)
    a = oracle.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, 1.0, 8.0)
    b = R.buddy_check(ob, x, y, q0, elev, g, NAMES[var], pressure, dx, 1.0, 8.0)
    for nm, u, v in zip(("qc", "aob", "err", "difmax", "buddy_num_final"), a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)
    assert np.count_nonzero(a[0] & (1 << 17)) > 0


@pytest.mark.parametrize("var,pressure", [(1, 300.0), (2, 850.0), (3, 1001.0), (3, 850.0), (3, 400.0), (4, 700.0),
                                          (5, 1001.0), (6, 1001.0), (7, 500.0)])
def test_error_max(oracle, var, pressure):
    """qc1_module.F90:102-249: factors per variable / level / local time / elevation / dewpoint band."""
    rng = np.random.default_rng(var * 31 + int(pressure))
    iew, jns, n = 60, 50, 300
    base = {1: 10.0, 2: 10.0, 3: 270.0, 4: 60.0, 5: 101000.0, 6: 101000.0, 7: 240.0}[var]
    spread = {5: 400.0, 6: 400.0, 7: 25.0}.get(var, 6.0)
    g = wavy(iew, jns, spread / 3, base)
    x = rng.uniform(2.0, iew - 1.0, n).astype(np.float32); y = rng.uniform(2.0, jns - 1.0, n).astype(np.float32)
    ob = (base + rng.normal(0, spread, n)).astype(np.float32)
    elev = rng.uniform(-500, 3500, n).astype(np.float32)
    lt = (rng.integers(0, 24, n) * 100).astype(np.int32)
    q0 = (rng.integers(0, 2, n) * (1 << 14) + rng.integers(0, 2, n) * (1 << 16)).astype(np.int32)
    maxd = {5: 300.0, 6: 300.0, 7: 12.0}.get(var, 5.0)
    a = oracle.error_max(ob, x, y, elev, q0, g, var, maxd, pressure, lt)
    b = R.error_max(ob, x, y, elev, q0, g, NAMES[var], maxd, pressure, lt)
    for nm, u, v in zip(("qc", "aob", "err", "thr"), a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)
    assert 0 < np.count_nonzero(a[0] != q0) < n


def test_ob_density_and_obs_distance(oracle):
    """qc0_module.F90:81-232: the window sweep and the brute-force nearest-box search as written."""
    rng = np.random.default_rng(12)
    iew, jns, n = 34, 27, 25
    x = rng.uniform(1.0, iew - 0.5, n).astype(np.float32); y = rng.uniform(1.0, jns - 0.5, n).astype(np.float32)
    tb0 = rng.integers(0, 2, (iew, jns)).astype(np.float32)
    for gd in (60.0, 110.0):
        a = oracle.ob_density(x, y, gd, tb0); b = R.ob_density(x, y, gd, tb0)
        np.testing.assert_array_equal(a, b)
    tb = np.zeros((iew, jns), np.float32)
    tb[rng.integers(0, iew, 6), rng.integers(0, jns, 6)] = 1.0
    tb[3, 4] = 0.5
    for dx in (3.0, 12.5):
        np.testing.assert_array_equal(oracle.obs_distance(tb, dx), R.obs_distance(tb, dx))
    np.testing.assert_array_equal(oracle.obs_distance(np.zeros((iew, jns), np.float32), 9.0),
                                  R.obs_distance(np.zeros((iew, jns), np.float32), 9.0))


def test_local_time(oracle):
    rng = np.random.default_rng(3)
    lon = np.concatenate([rng.uniform(-180, 180, 400), [-180.0, 180.0, 0.0, -7.5, 7.49, -172.6, 14.999]]).astype(np.float32)
    for hhmm in (0, 600, 1200, 1830, 2359):
        np.testing.assert_array_equal(oracle.local_time(hhmm, lon), R.local_time(hhmm, lon))


def test_qc_consistency(oracle):
    """qc0_module.F90:371-404 with contains_2n as written (obs_sort_module.F90:4515-4560)."""
    from obsgrid_b200 import synth
    case = synth.small_case(iew=30, jns=28, kbu=3, n_sfc=200, n_snd=50, seed=2)
    rng = np.random.default_rng(6)
    bits = np.array([1 << 14, 1 << 16, 1 << 17, 1 << 3, 1 << 11], np.int32)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh"):
        a = getattr(case, nm)
        a[...] = (rng.integers(0, 2, (5,) + a.shape).astype(np.int32) * bits[:, None, None]).sum(axis=0)
    want = [R.qc_consistency(case.qc_u[k], case.qc_v[k], case.qc_t[k], case.qc_rh[k]) for k in range(case.kbu)]
    t0 = case.qc_t.copy()
    oracle.qc_consistency(case)
    for k in range(case.kbu):
        np.testing.assert_array_equal(case.qc_u[k], want[k][0]); np.testing.assert_array_equal(case.qc_v[k], want[k][1])
        np.testing.assert_array_equal(case.qc_rh[k], want[k][2])
    np.testing.assert_array_equal(case.qc_t, t0)
    # contains_2n is a bit test on sums of distinct powers of two below 2**19
    for q in rng.integers(0, 1 << 19, 2000):
        for b in (14, 16, 17):
            assert R.contains_2n(int(q), 1 << b) == bool(int(q) >> b & 1)


@pytest.mark.parametrize("radii,mult,pressure", [((10, 8, 6, 4), 1.0, 1001.0), ((10, 8, 6, 4), 0.8, 1001.0),
                                                 ((10, 8, 6, 4), 0.5, 1001.0), ((10, 8, 6, 4), 0.5, 850.0),
                                                 ((200, 150, 90, 20), 0.75, 1001.0), ((5, 4), 0.8, 1001.0),
                                                 ((12, 9, 6, 4, 2), 1.0, 700.0)])
def test_scan_radius_logic(oracle, radii, mult, pressure):
    """proc_oa.F90:641-707 through ogo_oa_slab: scans done and the STOP case."""
    want, stop = R.scan_radii(list(radii) + [-1], mult, pressure)
    g = wavy(40, 36)
    x = np.array([10.5, 20.25], np.float32); y = np.array([12.5, 18.75], np.float32)
    ob = np.array([281.0, 279.0], np.float32)
    out, rc, ns = oracle.oa_slab(g, 3, pressure, ob, x, y, radii, 12.0, sfc_mult=mult)
    assert ns == len(want) and (rc != 0) == stop


@pytest.mark.parametrize("crsdot", [0, 1])
def test_smoothers(oracle, crsdot):
    """oa_module.F90:875-1006: interior stencil, the four edges, corners untouched, pass order."""
    g = (wavy(23, 17, 5.0, 0.0) + np.random.default_rng(1).normal(0, 1, (17, 23))).astype(np.float32)
    for passes in (1, 3):
        np.testing.assert_array_equal(oracle.smooth_5(g, passes, crsdot), R.smooth_5(g, passes, crsdot))
        np.testing.assert_array_equal(oracle.smoother_desmoother(g, passes, crsdot), R.smoother_desmoother(g, passes, crsdot))


def test_fg_t_at_point(oracle):
    """get_fg_at_point_module.F90:36-114: trapping levels, extrapolation below level 2, missing above
    the top and outside the domain."""
    from oracle.ogoracle import lib, ptr
    rng = np.random.default_rng(9)
    kbu, iew, jns = 7, 12, 10
    hlev = np.array([0.0, 100.0, 800.0, 1500.0, 3000.0, 5500.0, 9000.0], np.float32)
    h = (hlev[:, None, None] + rng.normal(0, 20, (kbu, iew, jns))).astype(np.float32)
    t = (290.0 - 0.0065 * h + rng.normal(0, 1, (kbu, iew, jns))).astype(np.float32)
    L = lib()
    n_missing = n_extrap = 0
    for it in range(400):
        x = float(rng.uniform(0.5, iew - 0.5)); y = float(rng.uniform(0.5, jns - 0.5))
        z = float(rng.uniform(-200, 9500) if it % 2 else rng.uniform(-200, 150))
        a = np.float32(L.ogo_fg_t_at_point(ptr(t), ptr(h), jns, iew, kbu, x, y, z))
        b = R.fg_t_at_point(t, h, x, y, z)
        assert a == b, (x, y, z, a, b)
        n_missing += a == np.float32(-888888.0); n_extrap += (z < 80.0 and a != np.float32(-888888.0))
    assert n_missing > 5 and n_extrap > 5
