"""Known-answer tests that pin the CPU oracle (oracle/og_oracle.c).

The reference ships no tests or golden vectors and cannot be compiled in this image
(Fortran + netCDF absent), so these hand-derived cases -- every expected value below is worked
out from the reference source lines cited, in explicit float32 arithmetic -- are the pin.
"""
import numpy as np
import pytest

f = np.float32


def flat_field(iew, jns, value=0.0):
    return np.full((jns, iew), value, np.float32)


def w_of(R, d2):
    R2 = f(R * R)
    return f(f(R2 - f(d2)) / f(R2 + f(d2)))


# ---- bilinear + innovation: proc_oa.F90:520-530 ------------------------------------------
def test_bilinear_plane_is_exact(oracle):
    iew, jns = 12, 9
    i = np.arange(1, iew + 1, dtype=np.float32)[None, :]
    j = np.arange(1, jns + 1, dtype=np.float32)[:, None]
    g = (2 * i + 4 * j).astype(np.float32)            # gridded(i,j) = 2i + 4j
    x = np.array([2.0, 3.5, 10.25, 11.0], np.float32)
    y = np.array([2.0, 4.25, 7.75, 8.0], np.float32)
    ob = np.array([100, 200, 300, 400], np.float32)
    diff, fg = oracle.innovation(g, x, y, ob)
    np.testing.assert_array_equal(fg, 2 * x + 4 * y)   # dyadic fractions: exact in float32
    np.testing.assert_array_equal(diff, ob - (2 * x + 4 * y))
    d2, _ = oracle.innovation(g, x, y, ob, scale=100.0)   # PMSL: obs/100 - aob (:528)
    np.testing.assert_array_equal(d2, (ob / f(100)).astype(np.float32) - (2 * x + 4 * y))
    d3, _ = oracle.innovation(g, x, y, ob, use_first_guess=False)  # :551
    np.testing.assert_array_equal(d3, ob)


# ---- dist_not_uu_vv_rh: oa_module.F90:612-620 ----------------------------------------------
def test_cressman_single_ob_circle(oracle):
    iew = jns = 24
    g = flat_field(iew, jns, 5.0)
    # ob at x = y = 10.5 -> xob - 0.5 = 10: sits exactly on grid point (10,10)
    out = oracle.cressman([2.0], [10.5], [10.5], g, var=3, radius=3, dxd_km=30.0)
    pert = out - g
    # one ob: pert = (w*w*diff)/w
    for di, dj in [(0, 0), (1, 0), (0, 2), (1, 1), (2, 2), (-2, 1)]:
        w = w_of(3, di * di + dj * dj)
        exp = f(f(f(w * w) * f(2.0)) / w)
        assert pert[10 - 1 + dj, 10 - 1 + di] == f(f(5.0) + exp) - f(5.0) or \
            out[10 - 1 + dj, 10 - 1 + di] == f(f(5.0) + exp)
    assert out[9, 9] == f(7.0)                 # d2 = 0: w = 1, pert = diff
    assert out[9, 10] == f(f(5.0) + f(f(f(f(0.8) * f(0.8)) * f(2.0)) / f(0.8)))  # d2=1: w=8/10
    assert out[9, 12] == f(5.0)                # d2 = 9 is not < 9 (:616)
    assert out[12, 9] == f(5.0)
    # only the open disc changed
    jj, ii = np.nonzero(out != g)
    assert set(zip(ii + 1, jj + 1)) == {(i, j) for i in range(1, 25) for j in range(1, 25)
                                        if (i - 10) ** 2 + (j - 10) ** 2 < 9}


def test_cressman_two_obs_sum_in_index_order(oracle):
    iew = jns = 20
    g = flat_field(iew, jns, 0.0)
    x = [8.5, 10.5]; y = [8.5, 8.5]; d = [1.0, -3.0]
    out = oracle.cressman(d, x, y, g, var=5, radius=4, dxd_km=30.0)
    # grid point (9,8): d2 to ob1 = 1, to ob2 = 1
    w = w_of(4, 1)
    num = f(f(0.0) + f(f(w * w) * f(1.0)))
    num = f(num + f(f(w * w) * f(-3.0)))
    den = f(f(f(0.0) + w) + w)
    assert out[8 - 1, 9 - 1] == f(num / den)
    # no first guess: gridded = perturbation everywhere (:215), zero where no ob reaches
    g5 = flat_field(iew, jns, 5.0)
    out2 = oracle.cressman(d, x, y, g5, var=5, radius=4, dxd_km=30.0, use_first_guess=False)
    assert out2[8 - 1, 9 - 1] == f(num / den)
    assert out2[19, 19] == 0.0 and out2[0, 0] == 0.0


def test_cressman_edges(oracle):
    iew, jns = 16, 14
    g = flat_field(iew, jns, 1.0)
    # obs at x <= 1.5, y <= 1.5, x >= iew-0.5, y >= jns-0.5 are skipped (:134-140)
    x = [1.5, 8.0, 15.5, 8.0]; y = [7.0, 1.5, 7.0, 13.5]
    out = oracle.cressman([9, 9, 9, 9], x, y, g, var=3, radius=3, dxd_km=30.0)
    np.testing.assert_array_equal(out, g)
    # an ob just inside updates up to iew-1 / jns-1 only (window :153-154)
    out = oracle.cressman([4.0], [15.4], [13.4], g, var=3, radius=3, dxd_km=30.0)
    assert np.all(out[:, iew - 1] == 1.0) and np.all(out[jns - 1, :] == 1.0)
    assert out[jns - 2, iew - 2] != 1.0


# ---- dist_uu_vv_rh shapes: oa_module.F90:309-313, 403, 468, 539-542 --------------------------
def test_cressman_ellipse_uniform_flow(oracle):
    iew = jns = 40
    g = flat_field(iew, jns, 0.0)
    ub = flat_field(iew, jns, 30.0); vb = flat_field(iew, jns, 0.0)   # straight westerly
    # p = 300 hPa -> vcrit = 15 (:309); speed 30 >= vcrit; v_ob = 0 -> radius 1000 (:388-390)
    # 1000 >= 3*R*dxd = 3*4*30 = 360 -> ellipse, elongation = 1 + 0.7778/15*30 (:479)
    oracle.reset_stats()
    out = oracle.cressman([1.0], [20.0], [20.0], g, var=1, radius=4, dxd_km=30.0,
                          u_banana=ub, v_banana=vb, pressure=300.0)
    st = oracle.get_stats()
    assert (st["n_circle"], st["n_ellipse"], st["n_banana"]) == (0, 1, 0)
    elong = f(f(1.0) + f(f(f(0.7778) / f(15.0)) * f(30.0)))
    # along-stream point (i - xob = 5, j = yob): x = 30*5/30 = 5, y = 0, d2 = 25/elong < 16
    d2 = f(f(f(5.0) * f(5.0)) / elong)
    assert d2 < 16
    w = w_of(4, d2)
    assert out[20 - 1, 25 - 1] == f(f(f(w * w) * f(1.0)) / w)
    # cross-stream the reach is the plain radius: (i = xob, j - yob = 4): d2 = 16, not < 16
    assert out[24 - 1, 20 - 1] == 0.0
    assert out[23 - 1, 20 - 1] != 0.0
    # the ellipse is centred on (xob, yob), not on (xob-0.5, yob-0.5) (:539 vs :502)
    assert out[20 - 1, 15 - 1] == out[20 - 1, 25 - 1]
    # slow flow at the same level -> circle (:403)
    oracle.reset_stats()
    oracle.cressman([1.0], [20.0], [20.0], g, var=1, radius=4, dxd_km=30.0,
                    u_banana=flat_field(iew, jns, 10.0), v_banana=vb, pressure=300.0)
    assert oracle.get_stats()["n_circle"] == 1
    # vcrit = 25 - 0.02 p at and below 500 hPa: 1000 hPa -> 5 m/s
    oracle.reset_stats()
    oracle.cressman([1.0], [20.0], [20.0], g, var=2, radius=4, dxd_km=30.0,
                    u_banana=flat_field(iew, jns, 6.0), v_banana=vb, pressure=1000.0)
    assert oracle.get_stats()["n_ellipse"] == 1


def test_cressman_banana_curved_flow(oracle):
    # solid-body rotation about (40,40): curved streamlines -> banana (:410)
    iew = jns = 90
    i = np.arange(1, iew + 1, dtype=np.float64)[None, :]
    j = np.arange(1, jns + 1, dtype=np.float64)[:, None]
    om = 2.0
    ub = (-om * (j - 40.0) + 0 * i).astype(np.float32)
    vb = (om * (i - 40.0) + 0 * j).astype(np.float32)
    g = flat_field(iew, jns, 0.0)
    oracle.reset_stats()
    out = oracle.cressman([1.0], [51.0], [51.0], g, var=1, radius=8, dxd_km=30.0,
                          u_banana=ub, v_banana=vb, pressure=300.0)
    assert oracle.get_stats()["n_banana"] == 1
    jj, ii = np.nonzero(out)
    # elongated: more grid points than the plain circle of the same radius reaches
    circ = oracle.cressman([1.0], [51.0], [51.0], g, var=3, radius=8, dxd_km=30.0)
    assert ii.size > np.count_nonzero(circ)
    # and never leaves the window of oa_module.F90:151-154 (NINT(50.5) = 51 -> 51 +- 33)
    assert ii.min() + 1 >= 51 - 33 and ii.max() + 1 <= min(51 + 33, iew - 1)
    assert jj.min() + 1 >= 51 - 33 and jj.max() + 1 <= min(51 + 33, jns - 1)
    # the flow at the ob is along (-1,+1): the influence region is longer along that diagonal
    along = ((ii - 50) - (jj - 50)) / np.sqrt(2.0)
    across = ((ii - 50) + (jj - 50)) / np.sqrt(2.0)
    assert np.ptp(along) > np.ptp(across)


def test_rh_scaling(oracle):
    # scale_cressman_rh_decreases: negative RH innovations are scaled by
    # min(1, gridded(i,j)/fg_value) (:563)
    iew = jns = 24
    g = flat_field(iew, jns, 80.0); g[:, 12:] = 40.0
    z = flat_field(iew, jns, 0.0)
    kw = dict(var=4, radius=4, dxd_km=30.0, u_banana=z, v_banana=z, pressure=700.0)
    plain = oracle.cressman([-10.0], [10.5], [10.5], g, fg_value=[80.0], scale_rh=False, **kw)
    scal = oracle.cressman([-10.0], [10.5], [10.5], g, fg_value=[80.0], scale_rh=True, **kw)
    assert plain[9, 9] == scal[9, 9] == f(70.0)          # gridded/fg = 1 there
    w = w_of(4, 9)                                       # (13,10): d2 = 9, gridded = 40
    sf = f(f(40.0) / f(80.0))
    assert scal[9, 12] == f(f(40.0) + f(f(f(f(w * w) * f(-10.0)) * sf) / w))
    assert plain[9, 12] == f(f(40.0) + f(f(f(w * w) * f(-10.0)) / w))
    pos = oracle.cressman([10.0], [10.5], [10.5], g, fg_value=[80.0], scale_rh=True, **kw)
    assert pos[9, 12] == f(f(40.0) + f(f(f(w * w) * f(10.0)) / w))   # positive: unscaled


# ---- multi-scan radius logic: proc_oa.F90:663-707 -------------------------------------------
def test_scan_radius_logic(oracle):
    iew = jns = 64
    g = flat_field(iew, jns, 0.0)
    ob = dict(obs_value=[1.0], xob=[30.5], yob=[30.5])
    # upper level: all three scans run regardless of sfc_mult
    _, rc, ns = oracle.oa_slab(g, 3, 500.0, radii=[30, 20, 10], dxd_km=3.0, sfc_mult=0.4, **ob)
    assert (rc, ns) == (0, 3)
    # surface, mult 0.4: 12, 8, 4.0 -> third is below 4.5 cells: stop after two scans (:694-704)
    _, rc, ns = oracle.oa_slab(g, 3, 1001.0, radii=[30, 20, 10], dxd_km=3.0, sfc_mult=0.4, **ob)
    assert (rc, ns) == (0, 2)
    # scan 1 below the floor: the reference STOPs (:695-703)
    _, rc, ns = oracle.oa_slab(g, 3, 1001.0, radii=[10, 5], dxd_km=3.0, sfc_mult=0.4, **ob)
    assert rc == -5 and ns == 0
    # list ends at the first negative entry (:643)
    _, rc, ns = oracle.oa_slab(g, 3, 1001.0, radii=[6, 4, -1, 3], dxd_km=3.0, **ob)
    assert ns == 2
    # RH is clamped to [0,100] after the scans (:803)
    out, _, _ = oracle.oa_slab(flat_field(iew, jns, 95.0), 4, 700.0, obs_value=[130.0],
                               xob=[30.5], yob=[30.5], radii=[4], dxd_km=3.0,
                               u_banana=g, v_banana=g)
    assert out.max() == 100.0 and out[29, 29] == 100.0


def test_second_scan_uses_reinterpolated_innovations(oracle):
    # scan 2 analyses ob - bilinear(field after scan 1) (:751-763).  Note the reference's
    # half-cell quirk: the Cressman circle is centred on (xob-.5, yob-.5) but the bilinear
    # first guess is taken at (xob, yob), so an ob "on" grid point (12,12) is not fitted exactly.
    iew = jns = 32
    g = flat_field(iew, jns, 0.0)
    one, _, _ = oracle.oa_slab(g, 3, 500.0, [4.0], [12.5], [12.5], radii=[3], dxd_km=30.0)
    two, _, _ = oracle.oa_slab(g, 3, 500.0, [4.0], [12.5], [12.5], radii=[3, 3], dxd_km=30.0)
    assert one[11, 11] == 4.0
    a = f(f(0.5) * f(f(f(0.5) * one[11, 11]) + f(f(0.5) * one[12, 11]))) + \
        f(f(0.5) * f(f(f(0.5) * one[11, 12]) + f(f(0.5) * one[12, 12])))
    diff2 = f(f(4.0) - f(a))
    assert diff2 != 0
    np.testing.assert_array_equal(
        two, oracle.cressman([diff2], [12.5], [12.5], one, var=3, radius=3, dxd_km=30.0))


# ---- smoothers: oa_module.F90:875-1006 -------------------------------------------------------
def test_smoothers(oracle):
    c = flat_field(12, 10, 3.0)
    np.testing.assert_array_equal(oracle.smooth_5(c, 3), c)
    d = flat_field(12, 10, 0.0); d[5, 6] = 8.0       # field(7,6) = 8
    s = oracle.smooth_5(d, 1)
    assert s[5, 6] == 4.0 and s[5, 7] == 1.0 and s[4, 6] == 1.0 and s[6, 6] == 1.0 and s[5, 5] == 1.0
    assert s.sum() == 8.0
    # boundary rows use the 1-2-1 stencil; i = iew - crsdot is the last smoothed column
    e = flat_field(12, 10, 0.0); e[4, 0] = 4.0        # field(1,5)
    s = oracle.smooth_5(e, 1)
    assert s[4, 0] == 2.0 and s[5, 0] == 1.0 and s[3, 0] == 1.0
    sd = oracle.smoother_desmoother(c, 2)
    np.testing.assert_array_equal(sd, c)
    sd = oracle.smoother_desmoother(d, 1)
    assert sd[5, 6] != 8.0 and sd[0, 0] == 0.0


# ---- flag arithmetic: qc0_module.F90:27-31, 55-72; obs_sort_module.F90:4515 -----------------
def test_flag_arithmetic_is_set_bit(oracle):
    T = 1 << 16
    q = np.arange(0, 1 << 19, 977, dtype=np.int32)
    err = np.ones(q.size, np.float32); thr = np.zeros(q.size, np.float32)
    np.testing.assert_array_equal(oracle.modify_qc_flag(q, err, thr, T), q | T)
    # not flagged when |err| <= thr (strict >)
    np.testing.assert_array_equal(oracle.modify_qc_flag(q, err, err, T), q)
    for v in (0, 5, 1 << 14, (1 << 14) + 3, (1 << 17) + (1 << 14), (1 << 18) - 1):
        assert oracle.add_to_qc_flag(v, 1 << 14) == v | (1 << 14)
    L = oracle.lib()
    for v in list(range(0, 4096)) + list(range((1 << 19) - 4096, 1 << 19)) + list(q):
        for n in (14, 16, 17):
            assert bool(L.ogo_contains_2n(int(v), 1 << n)) == bool(int(v) & (1 << n))


def test_local_time(oracle):
    # qc0_module.F90:261-267: INT(hh + lon/15), +24 if negative, mod 24, *100
    lt = oracle.local_time(1200, [-98.0, 0.0, 179.9, -180.0, 15.0, -7.4])
    np.testing.assert_array_equal(lt, [500, 1200, 2300, 0, 1300, 1100])
    lt = oracle.local_time(100, [-98.0, -16.0, 14.9])
    # 1 - 6.53 = -5.53 -> INT = -5 -> 19 ; 1 - 1.07 = -0.07 -> INT = 0 (truncation) ; 1.99 -> 1
    np.testing.assert_array_equal(lt, [1900, 0, 100])


# ---- error_max: qc1_module.F90:108-172, 220-249 ---------------------------------------------
def test_error_max_factors(oracle):
    iew = jns = 10
    g = flat_field(iew, jns, 280.0)
    x = np.full(4, 5.5, np.float32); y = np.full(4, 5.5, np.float32); el = np.zeros(4, np.float32)
    ob = np.array([280 + 12.4, 280 + 12.6, 280 + 14.9, 280 + 15.1], np.float32)
    # surface TT: factor 1.25 outside 1100..2000 local time, 1.5 inside; max_error_t = 10
    lt_night = np.array([300, 300, 300, 300], np.int32)
    qc, aob, err, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g, 3, 10.0, 1001.0, lt_night)
    np.testing.assert_array_equal(thr, np.full(4, 12.5, np.float32))
    np.testing.assert_array_equal(qc, [0, 1 << 16, 1 << 16, 1 << 16])
    lt_day = np.array([1100, 2000, 1500, 1500], np.int32)
    qc, *_ = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g, 3, 10.0, 1001.0, lt_day)
    np.testing.assert_array_equal(qc, [0, 0, 0, 1 << 16])
    # upper TT: 0.85 for 700 < p <= 1000.1, else 0.65
    _, _, _, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g, 3, 10.0, 850.0, lt_day)
    assert thr[0] == f(f(0.85) * f(10.0))
    _, _, _, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g, 3, 10.0, 700.0, lt_day)
    assert thr[0] == f(f(0.65) * f(10.0))
    # winds: max(factor*maxerr, 0.15*aob) (:228); factor 1.5 at p <= 499.9, 1.25 to 1000.1, 1 sfc
    gu = flat_field(iew, jns, 200.0)
    _, _, _, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), gu, 1, 13.0, 300.0, lt_day)
    assert thr[0] == f(f(0.15) * f(200.0))           # 30 > 1.5*13
    _, _, _, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g * 0, 2, 13.0, 850.0, lt_day)
    assert thr[0] == f(f(1.25) * f(13.0))
    _, _, _, thr = oracle.error_max(ob, x, y, el, np.zeros(4, np.int32), g * 0, 2, 13.0, 1001.0, lt_day)
    assert thr[0] == f(13.0)
    # DEWPOINT bands on the OB value (:157-165)
    obd = np.array([260.0, 250.0, 230.0, 210.0], np.float32)
    _, _, _, thr = oracle.error_max(obd, x, y, el, np.zeros(4, np.int32), g, 7, 20.0, 500.0, lt_day)
    np.testing.assert_array_equal(thr, np.array([20, 25, 30, 35], np.float32))
    # PMSLPSFC: 1 + |elev|/2000 (:149)
    elv = np.array([0, 1000, -1000, 2000], np.float32)
    _, _, _, thr = oracle.error_max(ob, x, y, elv, np.zeros(4, np.int32), g, 6, 600.0, 1001.0, lt_day)
    np.testing.assert_array_equal(thr, np.array([600, 900, 900, 1200], np.float32))


# ---- buddy_check: qc2_module.F90:121-337 ----------------------------------------------------
def test_buddy_check_by_hand(oracle):
    iew = jns = 40
    g = flat_field(iew, jns, 0.0)
    el = np.zeros(5, np.float32)
    # dx = 100 km, upper level (range 300 km): neighbours within 3 cells
    x = np.array([10, 11, 12, 13, 30], np.float32); y = np.full(5, 10, np.float32)
    ob = np.array([1.0, 2.0, 30.0, 3.0, 5.0], np.float32)   # innovations (first guess 0)
    qc, aob, err, dm, bnf = oracle.buddy_check(ob, x, y, np.zeros(5, np.int32), el, g, 3, 500.0,
                                               100.0, 1.0, 8.0)
    # TT at 500 hPa: difmax = 8*0.8 (:161-162)
    d0 = f(f(8.0) * f(0.8))
    # ob 1 (x=10): buddies x=11,12,13 (diffs 2,30,3): avg 35/3; 30 > 1.25*6.4 and |30-avg| > 6.4
    #   -> removed: kob = 2, sum = 35-30 = 5, err = 1 - 2.5 = -1.5 ; two buddies -> difmax*1.2
    avg = f(f(f(f(0) + f(2)) + f(30)) + f(3)) / f(3)
    assert abs(30 - avg) > d0
    assert bnf[0] == 2 and err[0] == f(f(1.0) - f(f(5.0) / f(2)))
    assert dm[0] == f(f(1.2) * d0)
    # ob 3 (the outlier): buddies 1,2,3 -> avg 2, err = 28 -> flagged
    assert bnf[2] == 3 and err[2] == f(f(30.0) - f(f(6.0) / f(3)))
    assert qc[2] == 1 << 17
    # ob 5 is alone: err = 0 and no_buddies (2**14) (:315-322)
    assert bnf[4] == 0 and err[4] == 0 and qc[4] == 1 << 14
    assert qc[0] == 0 and qc[1] == 0 and qc[3] == 0
    # distance < 0.1 km is not a buddy (:236): co-located obs do not see each other
    x2 = np.array([10, 10, 11], np.float32)
    qc2, *_ , bnf2 = oracle.buddy_check(ob[:3], x2, y[:3], np.zeros(3, np.int32), el[:3], g, 3,
                                        500.0, 100.0, 1.0, 8.0)
    assert list(bnf2) == [0, 0, 2]
    # range classes (:121-127): 650 km at p <= 201, 500 km at the surface
    x3 = np.array([10, 14, 16], np.float32)
    for p, nb in ((500.0, 0), (1001.0, 1), (200.0, 2)):
        _, _, _, _, b3 = oracle.buddy_check(ob[:3], x3, y[:3], np.zeros(3, np.int32), el[:3], g, 5,
                                            p, 100.0, 1.0, 800.0)
        # ob 1 sees 14 (400 km) only beyond 300 km range, 16 (600 km) only at 650 km
        assert (b3[0] >= 2) == (nb >= 2)


def test_buddy_relaxation_is_sequential(oracle):
    """difmax(num) *= 1.2 when two buddies survive (:327) is visible to LATER stations only."""
    iew = jns = 60
    g = flat_field(iew, jns, 0.0)
    # chain on a line, dx = 100 km, range 300 km.  Station A (index 0) has exactly 2 buddies.
    # Its diff is tuned so that B's removal test passes with difmax but fails with 1.2*difmax.
    d0 = f(8.0)                                            # PMSL-like: difmax = buddy_difference
    x = np.array([10, 12, 13, 14.5, 15.5], np.float32); y = np.full(5, 10, np.float32)
    a_diff = f(11.5)        # 1.25*8 = 10 < 11.5 < 1.25*9.6 = 12
    # B (x=12) has buddies A, C, D: average (11.5-3-3)/3 = 1.83, |11.5-1.83| = 9.67 > 8:
    # A is removed from B's set with difmax(A) = 8 but kept with the relaxed 9.6
    ob = np.array([a_diff, 0.0, -3.0, -3.0, 0.5], np.float32)
    el = np.zeros(5, np.float32)
    fwd = oracle.buddy_check(ob, x, y, np.zeros(5, np.int32), el, g, 5, 500.0, 100.0, 1.0, float(d0))
    # reversed order: A comes last, so nobody sees its relaxed difmax
    r = slice(None, None, -1)
    rev = oracle.buddy_check(ob[r].copy(), x[r].copy(), y[r].copy(), np.zeros(5, np.int32), el, g,
                             5, 500.0, 100.0, 1.0, float(d0))
    assert fwd[4][0] == 2                       # A: buddies at 12 and 13 only
    # station at x=12 sees A (200 km): with the relaxed difmax A is kept, otherwise removed
    assert fwd[4][1] == 3 and rev[4][::-1][1] == 2
    assert fwd[2][1] != rev[2][::-1][1]


def test_dewpoint_formula(oracle):
    assert oracle.dewpoint(280.0, 100.0) == f(280.0)          # LOG(1) = 0
    lo = oracle.dewpoint(280.0, 0.0)
    assert lo == oracle.dewpoint(280.0, 0.00005)               # clamped at very_small_rh
    td = oracle.dewpoint(290.0, 50.0)
    ref = 1.0 / (1.0 / 290.0 - np.log(0.5) / 5418.12)
    assert abs(td - ref) < 1e-3


def test_fg_t_at_point(oracle):
    # get_fg_at_point_module.F90:83-113: linear in height between trapping levels, then bilinear
    jns, iew, kbu = 6, 7, 5
    t = np.zeros((kbu, iew, jns), np.float32); h = np.zeros_like(t)
    for k in range(kbu):
        h[k] = 1000.0 * k           # level k+1 at 1000k m
        t[k] = 300.0 - 6.0 * k
    L = oracle.lib()
    from obsgrid_b200._capi import ptr
    v = L.ogo_fg_t_at_point(ptr(t), ptr(h), jns, iew, kbu, 3.5, 2.25, 2500.0)
    assert v == f(300.0 - 6.0 * 2.5)
    # below level 2: standard-atmosphere extrapolation from level 2 (:87-90)
    v = L.ogo_fg_t_at_point(ptr(t), ptr(h), jns, iew, kbu, 3.5, 2.25, 500.0)
    assert v == f(f(294.0) - f(f(-0.0065) * f(500.0)))
    # above the top or outside the domain: missing
    assert L.ogo_fg_t_at_point(ptr(t), ptr(h), jns, iew, kbu, 3.5, 2.25, 9000.0) == f(-888888.0)
    assert L.ogo_fg_t_at_point(ptr(t), ptr(h), jns, iew, kbu, 0.5, 2.25, 2500.0) == f(-888888.0)


def test_ob_density_and_obs_distance_by_hand(oracle):
    """qc0_module.F90:81-232 on a case small enough to do by hand."""
    iew, jns = 12, 9
    tb = np.zeros((iew, jns), np.float32)               # Fortran tobbox(jns,iew)
    # one ob at (x,y) = (4.3, 5.6), grid_dist 100 km -> radius 2.5 cells, window INT() +- 4
    out = oracle.ob_density([4.3], [5.6], 100.0, tb)
    want = np.zeros_like(tb)
    for i in range(1, iew):            # i <= iew-1
        for j in range(1, jns):        # j <= jns-1
            if np.sqrt(np.float32((i - 4.3) ** 2 + (j - 5.6) ** 2)) < 2.5:
                want[i - 1, j - 1] = 1.0
    assert want.sum() == 19 or want.sum() == 20 or want.sum() == 21   # a disc of radius 2.5 cells
    np.testing.assert_array_equal(out, want)
    # accumulation: a second call adds on top
    out2 = oracle.ob_density([4.3, 11.9], [5.6, 8.9], 100.0, out)
    assert out2[3, 5] == 2.0 and out2[iew - 1].sum() == 0 and out2[:, jns - 1].sum() == 0
    # distance: boxes (i,j) = (2,3) and (10,7) observed, dx = 3
    tb = np.zeros((iew, jns), np.float32)
    tb[1, 2] = 1.0; tb[9, 6] = 0.6; tb[5, 5] = 0.5        # 0.5 is not > 0.5
    d = oracle.obs_distance(tb, 3.0)
    assert d[1, 2] == 0.0 and d[9, 6] == 0.0
    assert d[5, 5] == np.float32(np.sqrt(np.float32(16 + 1))) * np.float32(3.0)     # nearest: (10,7)
    assert d[0, 0] == np.float32(np.sqrt(np.float32(1 + 4))) * np.float32(3.0)
    # no observed box at all: the domain diagonal (in cells) times dx
    d0 = oracle.obs_distance(np.zeros((iew, jns), np.float32), 3.0)
    diag = np.sqrt(np.float32(np.float32(11 * 3.0) * 11 * 3.0 + np.float32(8 * 3.0) * 8 * 3.0)) / np.float32(3.0)
    assert np.all(d0 == np.float32(diag) * np.float32(3.0))


# ---- post-analysis epilogue -------------------------------------------------
def _epilogue_fields(kbu, iew, jns, seed=0):
    rng = np.random.default_rng(seed)
    mk = lambda lo, hi: rng.uniform(lo, hi, (kbu, iew, jns)).astype(np.float32)
    return rng, mk


def test_wind_increment_to_c_grid_by_hand(oracle):
    """final_analysis_module.F90:316-339 with first_guess_module.F90:59-91, restated with numpy
    slices exactly as the Fortran array expressions are written (arrays (kbu,iew,jns) here)."""
    kbu, iew, jns = 3, 7, 5
    rng, mk = _epilogue_fields(kbu, iew, jns, 3)
    u, uA, uC, v, vA, vC = (mk(-30, 30) for _ in range(6))
    gu, gv = oracle.wind_increment_to_c_grid(u, uA, uC, v, vA, vC)
    half = np.float32(0.5)
    pu = u - uA                                   # pertubation(dim1=jns, dim2=iew, dim3)
    du = np.zeros_like(pu)
    du[:, 1:iew - 1, :] = (pu[:, 0:iew - 2, :] + pu[:, 1:iew - 1, :]) * half     # dummy(:,2:dim2-1,:)
    du[:, 0, :] = pu[:, 0, :]; du[:, iew - 1, :] = pu[:, iew - 2, :]
    np.testing.assert_array_equal(gu, uC + du)
    pv = v - vA
    dv = np.zeros_like(pv)
    dv[:, :, 1:jns - 1] = (pv[:, :, 0:jns - 2] + pv[:, :, 1:jns - 1]) * half     # dummy(2:dim1-1,:,:)
    dv[:, :, 0] = pv[:, :, 0]; dv[:, :, jns - 1] = pv[:, :, jns - 2]
    np.testing.assert_array_equal(gv, vC + dv)
    # a single-point increment of 2 at A-grid (j,i) spreads as 1 + 1 onto the two staggered neighbours
    u0 = np.zeros((1, 6, 4), np.float32); uA0 = u0.copy(); u0[0, 2, 1] = 2.0
    gu, _ = oracle.wind_increment_to_c_grid(u0, uA0, np.zeros_like(u0), u0, uA0, np.zeros_like(u0))
    assert gu[0, 2, 1] == 1.0 and gu[0, 3, 1] == 1.0 and gu.sum() == 2.0


def test_pres_sfc_from_slp_by_hand(oracle):
    """final_analysis_module.F90:364-370."""
    f = np.float32
    pres = np.array([[95000.0, 88000.5], [101300.0, 70000.0]], np.float32)
    slpc = np.array([[101000.0, 100500.0], [101300.0, 99000.0]], np.float32)
    slpx = np.array([[101250.0, 100100.0], [101300.0, 99900.0]], np.float32)
    out = oracle.pres_sfc_from_slp(pres, slpc, slpx)
    want = pres + (pres / slpc) * (slpx - slpc)          # float32 ops in source order
    np.testing.assert_array_equal(out, want)
    assert out[1, 0] == f(101300.0)                      # no SLP increment, no change


def test_mixing_ratio_and_sfcprs_against_float64(oracle):
    """final_analysis_module.F90:1026-1201: the FP32 restatement against the same formulas in
    float64 (tolerance: FP32 rounding of a dozen operations), plus structural facts."""
    press = np.array([1001, 1000, 925, 850, 700, 500, 300], np.float32)
    kbu, iew, jns = len(press), 9, 8
    rng = np.random.default_rng(11)
    t = np.empty((kbu, iew, jns), np.float32); h = np.empty_like(t)
    for k, p in enumerate(press):
        t[k] = 288.0 * (min(p, 1000.0) / 1000.0) ** 0.19 + rng.normal(0, 1.5, (iew, jns))
        h[k] = 44330.0 * (1 - (min(p, 1000.0) / 1013.25) ** 0.1903) + rng.normal(0, 10, (iew, jns))
    rh = rng.uniform(-5, 110, (kbu, iew, jns)).astype(np.float32)
    terrain = rng.uniform(0, 3200, (iew, jns)).astype(np.float32)   # all four branches of :1142 / :1173
    terrain[0, 0] = 0.0; terrain[1, 1] = 5200.0
    slp = rng.uniform(99000, 103000, (iew, jns)).astype(np.float32)
    rc, rh1, qv, psfc = oracle.mixing_ratio(rh, t, h, terrain, slp, press)
    assert rc == 0
    # clamp on the cross points only; last row / column untouched
    np.testing.assert_array_equal(rh1[:, :-1, :-1], np.clip(rh[:, :-1, :-1], 1.0, 100.0))
    np.testing.assert_array_equal(rh1[:, -1, :], rh[:, -1, :]); np.testing.assert_array_equal(rh1[:, :, -1], rh[:, :, -1])
    assert np.all(qv[:, -1, :] == 0) and np.all(qv[:, :, -1] == 0) and np.all(psfc[-1, :] == 0) and np.all(psfc[:, -1] == 0)
    T, H, RH = t.astype(np.float64), h.astype(np.float64), rh1.astype(np.float64)
    TER, SLP, P = terrain.astype(np.float64), slp.astype(np.float64), press.astype(np.float64)
    es = 6.112 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
    qs = 0.622 * es / (P[:, None, None] - es)
    q64 = np.maximum(0.01 * RH * qs, 0.0)
    k850, k700, k500 = 3, 4, 5
    R, g, gamma = 287.0, 9.806, 6.5e-3
    ps0 = SLP * (SLP / 85000.0) ** (-TER / H[k850])
    p1 = np.where(ps0 > 95000, 85000.0, np.where(ps0 > 70000, ps0 - 10000.0, 50000.0))
    tv = lambda k: T[k] * (1 + 0.608 * q64[k])
    g78 = np.log(tv(k850) / tv(k700)) / np.log(850.0 / 700.0)
    g57 = np.log(tv(k700) / tv(k500)) / np.log(700.0 / 500.0)
    t1 = np.where(ps0 > 95000, tv(k850), np.where(ps0 > 85000, tv(k700) * (p1 / 70000.0) ** g78,
                  np.where(ps0 > 70000, tv(k500) * (p1 / 50000.0) ** g57, tv(k500))))
    tslv = t1 * (SLP / p1) ** (gamma * R / g)
    tsfc = tslv - gamma * TER
    ps64 = SLP * np.exp(-TER * g / (R / 2.0 * (tsfc + tslv)))
    assert {int(x) for x in np.unique(np.digitize(ps0[:-1, :-1], [70000, 85000, 95000]))} == {0, 1, 2, 3}
    np.testing.assert_allclose(psfc[:-1, :-1], ps64[:-1, :-1], rtol=3e-6)
    np.testing.assert_allclose(qv[1:, :-1, :-1], q64[1:, :-1, :-1], rtol=2e-5, atol=1e-9)
    qs0 = 0.622 * es[0] / (ps64 / 100.0 - es[0])
    np.testing.assert_allclose(qv[0, :-1, :-1], np.maximum(0.01 * RH[0] * qs0, 0)[:-1, :-1], rtol=2e-5, atol=1e-9)
    # no 850/700/500 hPa level: the reference STOPs, the restatement reports it
    rc, *_ = oracle.mixing_ratio(rh[:4], t[:4], h[:4], terrain, slp, np.array([1001, 1000, 925, 850], np.float32))
    assert rc != 0


def test_qc_time_fdda_is_qc_time_plus_density(oracle):
    """proc_qc.F90:349-351, :412-420: the FDDA variant leaves the flags of ogo_qc_time and adds
    tobbox = (sum over the six checked surface variables of ob_density) / 5, zeroed below 1."""
    from obsgrid_b200 import synth
    base = synth.small_case(iew=50, jns=44, kbu=4, n_sfc=30, n_snd=6, radii=(6, 4), seed=5, dx_m=60000.0)
    a, b = base.copy(), base.copy()
    oracle.qc_time(a)
    tb, od = oracle.qc_time_fdda(b, np.zeros((base.iew, base.jns), np.float32))
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm))
    want = np.zeros((base.iew, base.jns), np.float32)
    ok_xy = (base.xob >= 2) & (base.xob <= base.iew - 1) & (base.yob >= 2) & (base.yob <= base.jns - 1)
    miss = lambda v: np.abs(v - np.float32(-888888.0)) <= 1
    for ob in (base.ob_u[0], base.ob_v[0], base.ob_t[0], base.ob_rh[0], base.ob_slp):
        m = ok_xy & ~miss(ob)
        want = oracle.ob_density(base.xob[m], base.yob[m], float(base.dx_km), want)
    m = ok_xy & ~miss(base.ob_rh[0]) & ~miss(base.ob_t[0])          # DEWPOINT: the sixth surface slab
    want = oracle.ob_density(base.xob[m], base.yob[m], float(base.dx_km), want)
    want = want / np.float32(5.0); want[want < 1] = 0
    np.testing.assert_array_equal(tb, want)
    np.testing.assert_array_equal(od, oracle.obs_distance(want, float(base.dx_km)))
    assert (tb > 0).any() and (tb == 0).any()


def test_temporal_interp_by_hand(oracle):
    """tinterp_module.F90:98-132: integer weights, cross-point edge replication."""
    rng = np.random.default_rng(4)
    a = rng.normal(280, 5, (3, 6, 5)).astype(np.float32); b = rng.normal(282, 5, (3, 6, 5)).astype(np.float32)
    f = np.float32
    # FDDA hour 2 of a 6-hour bracket: w1 = 6-2 = 4, w2 = 2-0 = 2
    rc, o = oracle.temporal_interp(a, b, 2, 0, 6, cross=False)
    assert rc == 0
    np.testing.assert_array_equal(o, (f(4) * a + f(2) * b) / f(6))
    rc, o = oracle.temporal_interp(a, b, 2, 0, 6, cross=True)
    want = (f(4) * a + f(2) * b) / f(6)
    want[:, :-1, -1] = want[:, :-1, -2]         # t(jns,:iew-1,:) = t(jns-1,:iew-1,:)
    want[:, -1, :] = want[:, -2, :]             # t(:jns,iew,:)   = t(:jns,iew-1,:)
    np.testing.assert_array_equal(o, want)
    # both bracketing times at the FDDA time: the first time is used as is (:106-109)
    rc, o = oracle.temporal_interp(a, b, 3, 3, 3, cross=False)
    np.testing.assert_array_equal(o, a)
    # weights summing to zero: the reference STOPs
    rc, _ = oracle.temporal_interp(a, b, 5, 7, 3, cross=False)      # w1 = -2, w2 = -2 ... sum -4: fine
    assert rc == 0
    rc, _ = oracle.temporal_interp(a, b, 4, 6, 2, cross=False)      # w1 = -2, w2 = -2
    assert rc == 0
    rc, _ = oracle.temporal_interp(a, b, 4, 2, 2, cross=False)      # w1 = -2, w2 = 2 -> 0
    assert rc != 0
