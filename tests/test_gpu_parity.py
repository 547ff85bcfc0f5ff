"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): QC flags bit-identical; analysed fields within 1e-5 relative /
1e-3 absolute.  The kernels keep the reference's summation order, so everything that does not
go through atan2 (banana shape) or log (DEWPOINT) is in fact compared bit-for-bit here.
"""
import numpy as np
import pytest

from obsgrid_b200 import synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-3       # north_star tolerance for analysed fields


def rng_obs(rng, n, iew, jns, lo=2.0):
    x = rng.uniform(lo, iew - 1.0, n).astype(np.float32)
    y = rng.uniform(lo, jns - 1.0, n).astype(np.float32)
    return x, y


def wavy(iew, jns, a=10.0, b=280.0, seed=0):
    i = np.arange(iew)[None, :]; j = np.arange(jns)[:, None]
    return (b + a * np.sin(i / 7.0 + seed) * np.cos(j / 5.0)).astype(np.float32)


def test_fast_div_is_exact(ctx):
    """The FCHK-less division used for the Cressman weight equals IEEE division."""
    assert ctx.selftest_div(1 << 24, seed=7, rmax=100) == 0
    assert ctx.selftest_div(1 << 22, seed=99, rmax=4000) == 0
    assert ctx.selftest_div(1 << 22, seed=5, rmax=7999) == 0      # the largest radius og_cressman admits


@pytest.mark.parametrize("iew,jns,n", [(40, 36, 300), (131, 77, 5000), (16, 16, 1), (33, 18, 0)])
def test_innovation_bit_exact(ctx, oracle, iew, jns, n):
    rng = np.random.default_rng(iew * 1000 + n)
    g = wavy(iew, jns)
    x, y = rng_obs(rng, n, iew, jns)
    ob = rng.normal(280, 8, n).astype(np.float32)
    for scale, ufg in ((1.0, True), (100.0, True), (1.0, False)):
        d0, f0 = oracle.innovation(g, x, y, ob, scale, ufg)
        d1, f1 = ctx.innovation(g, x, y, ob, scale, ufg)
        np.testing.assert_array_equal(d0, d1)
        if ufg:
            np.testing.assert_array_equal(f0, f1)


@pytest.mark.parametrize("var", [3, 5])
@pytest.mark.parametrize("iew,jns,n,R", [(64, 48, 400, 5), (100, 100, 2000, 10), (37, 29, 60, 2),
                                         (200, 150, 9000, 12), (20, 20, 5, 30)])
def test_cressman_circle_bit_exact(ctx, oracle, var, iew, jns, n, R):
    rng = np.random.default_rng(iew + 7 * n + R)
    g = wavy(iew, jns)
    x, y = rng_obs(rng, n, iew, jns, lo=1.0)      # some obs fall in the skipped edge band
    d = rng.normal(0, 3, n).astype(np.float32)
    a = oracle.cressman(d, x, y, g, var, R, 12.0)
    b = ctx.cressman(d, x, y, g, var, R, 12.0)
    np.testing.assert_array_equal(a, b)
    a = oracle.cressman(d, x, y, g, var, R, 12.0, use_first_guess=False)
    b = ctx.cressman(d, x, y, g, var, R, 12.0, use_first_guess=False)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("var", [3, 1])
def test_cressman_crsdot_zero_and_odd_grids(ctx, oracle, var):
    """crsdot = 0 (dot-point variables update the last row / column too) and grids that are not
    a multiple of the 32-point cell."""
    rng = np.random.default_rng(var)
    for iew, jns in ((33, 31), (65, 97), (5, 7)):
        n = max(4, iew * jns // 20)
        g = wavy(iew, jns)
        ub, vb = wind_fields(iew, jns, jet=45.0, curl=False)
        x, y = rng_obs(rng, n, iew, jns, lo=1.0)
        d = rng.normal(0, 3, n).astype(np.float32)
        for crsdot in (0, 1):
            a = oracle.cressman(d, x, y, g, var, 4, 12.0, ub, vb, 300.0, crsdot=crsdot)
            b = ctx.cressman(d, x, y, g, var, 4, 12.0, ub, vb, 300.0, crsdot=crsdot)
            np.testing.assert_array_equal(a, b, err_msg=f"{iew}x{jns} crsdot={crsdot}")


def test_cressman_dense_cluster_overflows_staging(ctx, oracle):
    """More candidates per tile than the shared-memory staging holds (batched consumption)."""
    rng = np.random.default_rng(5)
    iew, jns, n = 80, 70, 12000
    g = wavy(iew, jns)
    x = np.clip(rng.normal(40, 3.0, n), 2, iew - 1).astype(np.float32)
    y = np.clip(rng.normal(35, 3.0, n), 2, jns - 1).astype(np.float32)
    d = rng.normal(0, 2, n).astype(np.float32)
    np.testing.assert_array_equal(oracle.cressman(d, x, y, g, 3, 9, 12.0),
                                  ctx.cressman(d, x, y, g, 3, 9, 12.0))


def test_spatially_sorted_obs_exercise_the_sorter_fallback(ctx, oracle):
    """Obs whose index follows their position (sorted by column band): the indices reaching a
    cell then sit in a few narrow ranges, the value buckets of the list sorter are lopsided and
    its bitonic fallback has to produce the reference's order."""
    rng = np.random.default_rng(77)
    iew, jns, n = 96, 64, 9000
    g = wavy(iew, jns)
    x, y = rng_obs(rng, n, iew, jns)
    order = np.lexsort((rng.random(n), (x // 6).astype(np.int64) % 4, (y // 16).astype(np.int64)))
    x, y = x[order], y[order]
    d = rng.normal(0, 3, n).astype(np.float32)
    np.testing.assert_array_equal(oracle.cressman(d, x, y, g, 3, 8, 12.0), ctx.cressman(d, x, y, g, 3, 8, 12.0))
    ob = (280 + rng.normal(0, 5, n)).astype(np.float32)
    z = np.zeros(n, np.float32)
    a = oracle.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 3, 850.0, 40.0, 1.0, 8.0)
    b = ctx.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 3, 850.0, 40.0, 1.0, 8.0)
    for nm, u, v in zip(("qc", "aob", "err", "difmax", "buddy_num_final"), a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)


def wind_fields(iew, jns, jet=40.0, curl=True):
    i = np.arange(1, iew + 1, dtype=np.float64)[None, :]
    j = np.arange(1, jns + 1, dtype=np.float64)[:, None]
    yc = jns / 2 + 0.15 * jns * np.sin(2 * np.pi * i / iew * 1.5)
    core = np.exp(-((j - yc) / (0.12 * jns)) ** 2)
    u = 3.0 + jet * core
    v = jet * core * (0.15 * jns * 2 * np.pi * 1.5 / iew) * np.cos(2 * np.pi * i / iew * 1.5) if curl \
        else 0 * u
    return u.astype(np.float32), (v + 0 * u).astype(np.float32)


@pytest.mark.parametrize("var", [1, 2, 4])
@pytest.mark.parametrize("pressure", [300.0, 850.0, 1001.0])
def test_cressman_anisotropic(ctx, oracle, var, pressure):
    rng = np.random.default_rng(int(pressure) + var)
    iew, jns, n, R = 120, 96, 1500, 6
    g = wavy(iew, jns, 20.0, 50.0)
    ub, vb = wind_fields(iew, jns)
    x, y = rng_obs(rng, n, iew, jns, lo=1.0)
    d = rng.normal(0, 5, n).astype(np.float32)
    fg = np.abs(rng.normal(50, 20, n)).astype(np.float32) + 1
    oracle.reset_stats()
    a = oracle.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    st = oracle.get_stats()
    b = ctx.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    assert st["n_ellipse"] > 0
    if st["n_banana"] == 0:
        np.testing.assert_array_equal(a, b)
    else:   # atan2f: libdevice vs glibc differ in the last place -> tolerance, not bits
        np.testing.assert_allclose(b, a, rtol=RTOL, atol=ATOL)
        assert np.mean(a != b) < 0.2


def test_cressman_ellipse_only_bit_exact(ctx, oracle):
    rng = np.random.default_rng(11)
    iew, jns, n, R = 90, 60, 800, 5
    g = wavy(iew, jns)
    ub, vb = wind_fields(iew, jns, jet=45.0, curl=False)      # straight flow: v = 0 -> ellipse
    x, y = rng_obs(rng, n, iew, jns)
    d = rng.normal(0, 5, n).astype(np.float32)
    oracle.reset_stats()
    a = oracle.cressman(d, x, y, g, 1, R, 12.0, ub, vb, 300.0)
    st = oracle.get_stats()
    assert st["n_banana"] == 0 and st["n_ellipse"] > 50 and st["n_circle"] > 50
    np.testing.assert_array_equal(a, ctx.cressman(d, x, y, g, 1, R, 12.0, ub, vb, 300.0))


@pytest.mark.parametrize("smooth_type,passes", [(1, 1), (1, 3), (2, 2)])
def test_cressman_smoothing_bit_exact(ctx, oracle, smooth_type, passes):
    rng = np.random.default_rng(3)
    iew, jns, n = 70, 55, 500
    g = wavy(iew, jns)
    x, y = rng_obs(rng, n, iew, jns)
    d = rng.normal(0, 3, n).astype(np.float32)
    a = oracle.cressman(d, x, y, g, 3, 6, 12.0, passes=passes, smooth_type=smooth_type)
    b = ctx.cressman(d, x, y, g, 3, 6, 12.0, passes=passes, smooth_type=smooth_type)
    np.testing.assert_array_equal(a, b)


def test_oa_slab_multi_scan(ctx, oracle):
    rng = np.random.default_rng(21)
    iew, jns, n = 150, 120, 3000
    g = wavy(iew, jns, 25.0, 60.0)
    ub, vb = wind_fields(iew, jns, curl=False)
    x, y = rng_obs(rng, n, iew, jns)
    for var, scale, p in ((3, 1.0, 500.0), (5, 100.0, 1001.0), (4, 1.0, 700.0)):
        ob = (rng.normal(60, 15, n) * scale).astype(np.float32)
        a, rca, nsa = oracle.oa_slab(g, var, p, ob, x, y, [10, 8, 6, 4], 12.0, ub, vb, scale,
                                     scale_rh=True)
        b, rcb, nsb = ctx.oa_slab(g, var, p, ob, x, y, [10, 8, 6, 4], 12.0, ub, vb, scale,
                                  scale_rh=True)
        assert (rca, nsa) == (rcb, nsb) == (0, 4)
        np.testing.assert_array_equal(a, b)
    # surface radius scaling and the 4.5-cell floor (proc_oa.F90:663-707)
    a, rca, nsa = oracle.oa_slab(g, 3, 1001.0, ob, x, y, [30, 20, 10], 3.0, sfc_mult=0.4)
    b, rcb, nsb = ctx.oa_slab(g, 3, 1001.0, ob, x, y, [30, 20, 10], 3.0, sfc_mult=0.4)
    assert (rca, nsa) == (rcb, nsb) == (0, 2)
    np.testing.assert_array_equal(a, b)
    a, rca, nsa = oracle.oa_slab(g, 3, 1001.0, ob, x, y, [10, 5], 3.0, sfc_mult=0.4)
    b, rcb, nsb = ctx.oa_slab(g, 3, 1001.0, ob, x, y, [10, 5], 3.0, sfc_mult=0.4)
    assert (rca, nsa) == (rcb, nsb) == (-5, 0)
    np.testing.assert_array_equal(a, b)


def test_fg_t_at_point_bit_exact(ctx, oracle):
    """GET_FG_T_AT_POINT: trapping levels, extrapolation below level 2, missing above
    the top and outside the domain."""
    import ctypes as C
    rng = np.random.default_rng(14)
    kbu, iew, jns, n = 9, 30, 26, 2000
    base = np.linspace(0.0, 12000.0, kbu, dtype=np.float32)[:, None, None]
    h = (base + rng.uniform(-150, 150, (kbu, iew, jns))).astype(np.float32)
    h = np.sort(h, axis=0)
    t = (288.0 - 0.0065 * h + rng.normal(0, 1.0, h.shape)).astype(np.float32)
    x = rng.uniform(0.2, iew - 0.2, n).astype(np.float32)
    y = rng.uniform(0.2, jns - 0.2, n).astype(np.float32)
    z = rng.uniform(-500.0, 13500.0, n).astype(np.float32)
    x[:10] = np.arange(1, 11); y[:10] = np.arange(1, 11)          # exactly on grid points
    got = ctx.fg_t_at_point(t, h, x, y, z)
    L = oracle.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    want = np.array([L.ogo_fg_t_at_point(p(t), p(h), jns, iew, kbu, float(x[k]), float(y[k]), float(z[k]))
                     for k in range(n)], np.float32)
    np.testing.assert_array_equal(got, want)
    assert 50 < np.count_nonzero(want == np.float32(-888888.0)) < n - 50


@pytest.mark.parametrize("iew,jns,n,gd", [(60, 45, 300, 30.0), (33, 70, 40, 12.0), (40, 40, 0, 30.0), (90, 64, 2500, 50.0)])
def test_ob_density_and_obs_distance_bit_exact(ctx, oracle, iew, jns, n, gd):
    """qc0_module.F90:81-232; the distance is an exact EDT on the device."""
    rng = np.random.default_rng(iew + n)
    x = rng.uniform(-3.0, iew + 3.0, n).astype(np.float32)      # some obs outside the domain
    y = rng.uniform(-3.0, jns + 3.0, n).astype(np.float32)
    tb0 = (rng.random((iew, jns)) < 0.02).astype(np.float32) * 3.0
    a = oracle.ob_density(x, y, gd, tb0)
    b = ctx.ob_density(x, y, gd, tb0)
    np.testing.assert_array_equal(a, b)
    # sparse occupancy for the distance: a handful of observed boxes, then none, then the density field
    for tb in ((rng.random((iew, jns)) < 0.004).astype(np.float32), np.zeros((iew, jns), np.float32),
               np.where(a / 5.0 < 1.0, 0.0, a / 5.0).astype(np.float32)):
        for dx in (gd, 1.0, 3000.0):
            np.testing.assert_array_equal(oracle.obs_distance(tb, dx), ctx.obs_distance(tb, dx))


def test_local_time_bit_exact(ctx, oracle):
    rng = np.random.default_rng(2)
    lon = rng.uniform(-180, 180, 5000).astype(np.float32)
    for hhmm in (0, 100, 1200, 2359):
        np.testing.assert_array_equal(oracle.local_time(hhmm, lon), ctx.local_time(hhmm, lon))


@pytest.mark.parametrize("var,pressure", [(1, 300.0), (2, 1001.0), (3, 1001.0), (3, 850.0), (3, 500.0),
                                          (4, 700.0), (5, 1001.0), (6, 1001.0), (7, 500.0)])
def test_error_max_bit_exact(ctx, oracle, var, pressure):
    rng = np.random.default_rng(var * 10 + int(pressure))
    iew, jns, n = 90, 70, 4000
    g = wavy(iew, jns, 12.0, 270.0 if var in (3, 7) else 20.0)
    x, y = rng_obs(rng, n, iew, jns)
    ob = (g.mean() + rng.normal(0, 12, n)).astype(np.float32)
    elev = rng.uniform(-500, 3000, n).astype(np.float32)
    lt = (rng.integers(0, 24, n) * 100).astype(np.int32)
    q0 = (rng.integers(0, 4, n) * (1 << 14)).astype(np.int32)
    a = oracle.error_max(ob, x, y, elev, q0, g, var, 10.0, pressure, lt)
    b = ctx.error_max(ob, x, y, elev, q0, g, var, 10.0, pressure, lt)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)
    assert 0 < np.count_nonzero(a[0] & (1 << 16)) < n


@pytest.mark.parametrize("var,pressure,n,dx", [(3, 1001.0, 3000, 30.0), (1, 300.0, 2500, 30.0),
                                               (5, 1001.0, 700, 12.0), (4, 150.0, 1200, 60.0),
                                               (7, 500.0, 300, 90.0), (6, 1001.0, 900, 30.0),
                                               (3, 500.0, 1, 30.0), (3, 500.0, 2, 30.0),
                                               (2, 700.0, 257, 150.0)])
def test_buddy_check_bit_exact(ctx, oracle, var, pressure, n, dx):
    rng = np.random.default_rng(var + n)
    iew, jns = 110, 90
    g = wavy(iew, jns, 6.0, 260.0 if var in (3, 7) else 10.0)
    x, y = rng_obs(rng, n, iew, jns)
    if n > 10:                       # co-located pairs: distance < 0.1 km is not a buddy
        x[1], y[1] = x[0], y[0]
    ob = (g.mean() + rng.normal(0, 5, n)).astype(np.float32)
    ob[rng.random(n) < 0.05] += 40.0
    elev = rng.uniform(-3000, 3000, n).astype(np.float32)
    q0 = (rng.integers(0, 2, n) * (1 << 16)).astype(np.int32)
    a = oracle.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, 1.0, 8.0)
    b = ctx.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, 1.0, 8.0)
    names = ("qc", "aob", "err", "difmax", "buddy_num_final")
    for nm, u, v in zip(names, a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)


def test_buddy_check_long_cell_lists(ctx, oracle):
    """Dense network on a coarse grid: every cell's candidate list is longer than the
    shared-memory sorter (8192), so the value-bucket sorter of og_bin.cu orders it; a Gaussian
    cluster makes the cells very unequal (one CTA per (cell, 128 targets) work item)."""
    rng = np.random.default_rng(21)
    iew, jns, n = 40, 40, 30000
    g = wavy(iew, jns, 6.0, 280.0)
    x, y = rng_obs(rng, n, iew, jns)
    k = n // 3
    x[:k] = np.clip(rng.normal(20, 2.0, k), 2, iew - 1).astype(np.float32)
    y[:k] = np.clip(rng.normal(22, 2.0, k), 2, jns - 1).astype(np.float32)
    p = rng.permutation(n); x, y = x[p], y[p]
    ob = (280 + rng.normal(0, 5, n)).astype(np.float32)
    ob[rng.random(n) < 0.03] += 30.0
    z = np.zeros(n, np.float32)
    ctx.reset_stats()
    a = oracle.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 3, 850.0, 50.0, 1.0, 8.0)
    b = ctx.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 3, 850.0, 50.0, 1.0, 8.0)
    for nm, u, v in zip(("qc", "aob", "err", "difmax", "buddy_num_final"), a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)
    st = ctx.stats()
    assert st["buddy_pairs_examined"] < st["buddy_pair_tests"]


def test_buddy_relaxation_fixpoint(ctx, oracle):
    """Sparse network where many obs keep exactly two buddies: the x1.2 relaxation of
    qc2_module.F90:327 feeds later stations; the GPU fix-point must equal the serial loop."""
    rng = np.random.default_rng(8)
    iew, jns, n = 400, 400, 1500
    g = np.zeros((jns, iew), np.float32)
    x, y = rng_obs(rng, n, iew, jns)
    ob = rng.normal(0, 6, n).astype(np.float32)
    ob[rng.random(n) < 0.3] += 11.0          # many diffs between 1.25*d and 1.25*1.2*d
    z = np.zeros(n, np.float32)
    ctx.reset_stats()
    a = oracle.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 5, 500.0, 30.0, 1.0, 8.0)
    b = ctx.buddy_check(ob, x, y, np.zeros(n, np.int32), z, g, 5, 500.0, 30.0, 1.0, 8.0)
    assert np.count_nonzero(a[4] == 2) > 20
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)
    assert ctx.stats()["buddy_fixpoint_iters"] >= 2


def test_qc_slab_fused(ctx, oracle):
    rng = np.random.default_rng(13)
    iew, jns, n = 100, 100, 2000
    g = wavy(iew, jns, 8.0, 285.0)
    x, y = rng_obs(rng, n, iew, jns)
    ob = (285 + rng.normal(0, 6, n)).astype(np.float32)
    ob[rng.random(n) < 0.03] -= 30
    elev = np.zeros(n, np.float32)
    lt = (rng.integers(0, 24, n) * 100).astype(np.int32)
    q0 = np.zeros(n, np.int32)
    qa, *_ = oracle.error_max(ob, x, y, elev, q0, g, 3, 10.0, 1001.0, lt)
    qa, *_ = oracle.buddy_check(ob, x, y, qa, elev, g, 3, 1001.0, 30.0, 1.0, 8.0)
    qb = ctx.qc_slab(g, 3, 1001.0, ob, x, y, elev, lt, q0, 10.0, 8.0, 1.0, 30.0)
    np.testing.assert_array_equal(qa, qb)
    assert np.count_nonzero(qa) > 0


def _compare_time(case_gpu, case_cpu, exact_fields=("t", "slp"), tol_fields=("u", "v", "rh")):
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(case_cpu, nm), getattr(case_gpu, nm), err_msg=nm)
    for nm in exact_fields:
        np.testing.assert_array_equal(getattr(case_cpu, nm), getattr(case_gpu, nm), err_msg=nm)
    for nm in tol_fields:
        np.testing.assert_allclose(getattr(case_gpu, nm), getattr(case_cpu, nm), rtol=RTOL,
                                   atol=ATOL, err_msg=nm)


def _rh_flag_mismatches(case_gpu, case_cpu):
    """DEWPOINT goes through LOG; the device restates glibc's logf (csrc/og_logf.h, tests/test_logf.py),
    so the RH / DEWPOINT flags are bit-identical like every other flag.  Returns the mismatching
    (k, r) list, which every test requires to be empty."""
    return np.argwhere(case_gpu.qc_rh != case_cpu.qc_rh)


@pytest.mark.parametrize("scale_rh,sfc_mult", [(False, 1.0), (True, 0.8)])
def test_time_small(ctx, oracle, scale_rh, sfc_mult):
    base = synth.small_case(iew=60, jns=52, kbu=6, n_sfc=500, n_snd=120, radii=(6, 4, 2),
                            scale_rh=scale_rh, sfc_mult=sfc_mult)
    cpu, gpu = base.copy(), base.copy()
    oracle.qc_time(cpu); ctx.proc_qc(gpu)
    assert len(_rh_flag_mismatches(gpu, cpu)) == 0
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm), err_msg=nm)
    assert np.count_nonzero(cpu.qc_t) > 0 and np.count_nonzero(cpu.qc_rh) > 0
    # each side analyses from its own flags
    assert oracle.oa_time(cpu) == ctx.proc_oa(gpu) == 0
    _compare_time(gpu, cpu)


@pytest.mark.parametrize("groups", [0, 1, 3, 5])
def test_fused_pipelined_time_equals_the_two_stage_calls(ctx, groups):
    """og_qc_oa_time (levels pipelined through the copy engines) == og_qc_time_levels followed
    by og_oa_time, bit for bit, for any grouping of the levels."""
    base = synth.small_case(iew=64, jns=48, kbu=9, n_sfc=500, n_snd=120, seed=5, radii=(6, 4, 2))
    a, b = base.copy(), base.copy()
    ctx.proc_qc(a); ctx.proc_oa(a)
    ctx.proc_qc_oa(b, pipeline_groups=groups)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)
    c = base.copy()
    ctx.proc_qc_oa(c, 2, 7, pipeline_groups=2)          # a level sub-range leaves the rest alone
    for nm in ("t", "u", "qc_t"):
        np.testing.assert_array_equal(getattr(c, nm)[2:7], getattr(a, nm)[2:7], err_msg=nm)
        np.testing.assert_array_equal(getattr(c, nm)[:2], getattr(base, nm)[:2], err_msg=nm)
        np.testing.assert_array_equal(getattr(c, nm)[7:], getattr(base, nm)[7:], err_msg=nm)


def test_async_time_slots_pipeline(ctx):
    """Several time periods in flight through og_time_upload/_compute/_download/_wait give the
    results of the blocking calls; inputs are left untouched when the output arrays differ."""
    times = [synth.small_case(iew=48, jns=40, kbu=6, n_sfc=300, n_snd=80, seed=30 + t, radii=(5, 3))
             for t in range(5)]
    want = [c.copy() for c in times]
    for c in want:
        ctx.proc_qc(c); ctx.proc_oa(c)
    pristine = [c.copy() for c in times]
    outs = [c.copy() for c in times]
    NS = 3
    ctx.time_upload(times[0], 0)
    for i, c in enumerate(times):
        if i + 1 < len(times):
            ctx.time_upload(times[i + 1], (i + 1) % NS)
        ctx.time_compute(c, i % NS)
        ctx.time_download(outs[i], i % NS)
    for s in range(NS):
        ctx.time_wait(s)
    for i in range(len(times)):
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
            np.testing.assert_array_equal(getattr(outs[i], nm), getattr(want[i], nm), err_msg=f"{i}:{nm}")
            np.testing.assert_array_equal(getattr(times[i], nm), getattr(pristine[i], nm))


@pytest.mark.parametrize("with_td", [False, True])
def test_sparse_observation_rows_equal_the_dense_slot_path(ctx, with_td):
    """og_time_upload_csr / og_time_download_csr (the observation rows as per-level lists of the reports
    that have a measurement there) == og_time_upload / og_time_download, for level sub-ranges too."""
    times = [synth.small_case(iew=48, jns=40, kbu=6, n_sfc=300, n_snd=80, seed=40 + t, radii=(5, 3)) for t in range(3)]
    if with_td:
        rng = np.random.default_rng(1)
        for c in times:
            c.ob_td = np.where(np.abs(c.ob_t + 888888.0) > 1, c.ob_t - rng.uniform(-2, 10, c.ob_t.shape), -888888.0).astype(np.float32)
            c.qc_td = np.zeros_like(c.qc_t)
            c.bogus = (rng.random(c.nrep) < 0.05).astype(np.int32)
    want = [c.copy() for c in times]
    for c in want:
        ctx.proc_qc(c); ctx.proc_oa(c)
    for (k0, k1) in ((0, 6), (2, 5)):
        for i, c in enumerate(times):
            csr_in, out = synth.CsrObs(c), c.copy()
            csr_out = synth.CsrObs(c)
            slot = i % 2
            ctx.time_upload_csr(c, csr_in, slot, k0, k1)
            ctx.time_compute(c, slot, k0, k1)
            ctx.time_download_csr(out, csr_out, slot, k0, k1)
            ctx.time_wait(slot)
            csr_out.scatter_flags(out)
            names = ("qc_u", "qc_v", "qc_t", "qc_rh", "t", "u", "v", "rh") + (("qc_td", "ob_td") if with_td else ())
            for nm in names:
                np.testing.assert_array_equal(getattr(out, nm)[k0:k1], getattr(want[i], nm)[k0:k1], err_msg=f"{i}:{nm}")
                np.testing.assert_array_equal(getattr(out, nm)[:k0], getattr(c, nm)[:k0], err_msg=f"{i}:{nm} below the range")
            if k0 == 0:
                np.testing.assert_array_equal(out.qc_slp, want[i].qc_slp)
                np.testing.assert_array_equal(out.slp, want[i].slp)
            assert csr_in.nnz < c.kbu * c.nrep // 2


def test_time_level_sharding_is_exact(ctx):
    """Levels are independent: ranges and variable masks compose to the full run."""
    base = synth.small_case(iew=50, jns=44, kbu=7, n_sfc=300, n_snd=100, radii=(5, 3))
    full, parts = base.copy(), base.copy()
    ctx.proc_qc(full); ctx.proc_oa(full)
    for k0, k1 in ((3, 7), (0, 1), (1, 3)):
        ctx.proc_qc(parts, k0, k1)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(full, nm), getattr(parts, nm), err_msg=nm)
    for k0, k1 in ((4, 7), (0, 2), (2, 4)):
        ctx.proc_oa(parts, k0, k1)
    for nm in ("t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(getattr(full, nm), getattr(parts, nm), err_msg=nm)


@pytest.mark.parametrize("scale_rh,sfc_mult", [(False, 1.0), (True, 0.8)])
def test_per_slab_drop_ins_equal_the_batched_time(ctx, scale_rh, sfc_mult):
    """The reference's own slab loops calling og_error_max / og_buddy_check / og_innovation / og_cressman
    one slab and stage at a time (obsgrid_b200/perslab.py: what an unmodified proc_qc / proc_oa does
    over the Fortran wrappers) give the batched per-time result bit for bit."""
    from obsgrid_b200 import perslab
    base = synth.small_case(iew=60, jns=52, kbu=5, n_sfc=400, n_snd=100, radii=(6, 4, 2), scale_rh=scale_rh,
                            sfc_mult=sfc_mult)
    a, b = base.copy(), base.copy()
    ctx.proc_qc(a); ctx.proc_oa(a)
    calls = perslab.proc_qc_per_slab(ctx, b) + perslab.proc_oa_per_slab(ctx, b)
    assert calls > 100
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)


def test_time_cfg1(ctx, oracle):
    """BASELINE.json configs[0]: 100x100, 27 levels, 2k reports, 4 scans + qc1/qc2."""
    base = synth.make_case(1)
    cpu, gpu = base.copy(), base.copy()
    oracle.qc_time(cpu, threads=8); ctx.proc_qc(gpu)
    assert len(_rh_flag_mismatches(gpu, cpu)) == 0
    oracle.reset_stats(); ctx.reset_stats(); ctx.set_counting(True)
    assert oracle.oa_time(cpu, threads=8) == ctx.proc_oa(gpu) == 0
    ctx.set_counting(False)
    _compare_time(gpu, cpu)
    # the GPU counts exactly the reference's contributing updates where shapes are exact
    so, sg = oracle.get_stats(), ctx.stats()
    assert abs(sg["cressman_updates"] - so["cressman_updates"]) <= 1e-4 * so["cressman_updates"]
    assert sg["cressman_candidates"] < so["cressman_candidates"]


def test_empty_and_degenerate_inputs(ctx, oracle):
    g = wavy(30, 20)
    e = np.zeros(0, np.float32)
    np.testing.assert_array_equal(ctx.cressman(e, e, e, g, 3, 5, 12.0), g)
    z = np.zeros((20, 30), np.float32)
    np.testing.assert_array_equal(ctx.cressman(e, e, e, g, 3, 5, 12.0, use_first_guess=False), z)
    q = ctx.qc_slab(g, 3, 500.0, e, e, e, e, np.zeros(0, np.int32), np.zeros(0, np.int32), 10, 8, 1, 30)
    assert q.size == 0
    case = synth.small_case(n_sfc=0, n_snd=0)
    ref = case.copy()
    ctx.proc_qc(case); ctx.proc_oa(case)
    np.testing.assert_array_equal(case.t, ref.t)
    # RH is clamped even without obs (mixing_ratio clamp + clean_rh)
    assert case.rh.min() >= 0 and case.rh.max() <= 100


def test_obs_outside_the_slab_are_refused(ctx):
    """The per-slab entry points interpolate at every ob: coordinates outside [1, iew) x [1, jns) or
    NaN would read out of bounds; they come back as OG_ERR_INVALID and the context stays usable."""
    from obsgrid_b200._capi import OgError
    g = wavy(30, 20)
    one = np.ones(1, np.float32)
    for x, y in ((0.5, 5.0), (30.0, 5.0), (5.0, 20.0), (np.nan, 5.0), (5.0, np.inf), (-3.0, 5.0)):
        xs, ys = one * np.float32(x), one * np.float32(y)
        with pytest.raises(OgError):
            ctx.cressman(one, xs, ys, g, 3, 5, 12.0)
        with pytest.raises(OgError):
            ctx.innovation(g, xs, ys, one)
        with pytest.raises(OgError):
            ctx.qc_slab(g, 3, 500.0, one, xs, ys, one, np.zeros(1, np.int32), np.zeros(1, np.int32), 10, 8, 1, 30)
    out = ctx.cressman(one, one * 5, one * 5, g, 3, 5, 12.0)      # the context is not poisoned
    assert np.any(out != g)


def test_radius_zero_scan_is_a_no_op(ctx, oracle):
    """radius_influence = 0: no squared distance is below 0, the reference updates nothing
    (oa_module.F90:553 / :616); every shape, including a banana whose centre is a grid point."""
    rng = np.random.default_rng(3)
    iew, jns, n = 60, 50, 400
    g = wavy(iew, jns, 20.0, 50.0)
    ub, vb = wind_fields(iew, jns)
    x, y = rng_obs(rng, n, iew, jns)
    d = rng.normal(0, 5, n).astype(np.float32)
    for var, p in ((3, 500.0), (1, 300.0), (4, 1001.0)):
        a = oracle.cressman(d, x, y, g, var, 0, 12.0, ub, vb, p)
        b = ctx.cressman(d, x, y, g, var, 0, 12.0, ub, vb, p)
        np.testing.assert_array_equal(a, g)
        np.testing.assert_array_equal(b, g)


def test_bad_arguments_are_reported(ctx):
    from obsgrid_b200._capi import OgError
    g = wavy(30, 20)
    with pytest.raises(OgError):
        ctx.cressman([1.0], [5.0], [5.0], g, 1, 5, 12.0)       # UU without banana winds
    with pytest.raises(OgError):
        ctx.cressman([1.0], [5.0], [5.0], g, 9, 5, 12.0)       # unknown variable


# ---- post-analysis epilogue -------------------------------------------------
@pytest.mark.parametrize("kbu,iew,jns", [(3, 7, 5), (27, 100, 100), (1, 2, 2), (5, 333, 257)])
def test_wind_increment_to_c_grid_bit_exact(ctx, oracle, kbu, iew, jns):
    """final_analysis_module.F90:316-339 + first_guess_module.F90:59-91."""
    rng = np.random.default_rng(kbu * 1000 + iew)
    mk = lambda: rng.uniform(-40, 40, (kbu, iew, jns)).astype(np.float32)
    u, uA, uC, v, vA, vC = (mk() for _ in range(6))
    a = oracle.wind_increment_to_c_grid(u, uA, uC, v, vA, vC)
    b = ctx.wind_increment_to_c_grid(u, uA, uC, v, vA, vC)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_pres_sfc_from_slp_bit_exact(ctx, oracle):
    """final_analysis_module.F90:364-370."""
    rng = np.random.default_rng(5)
    iew, jns = 211, 173
    pres = rng.uniform(55000, 103000, (iew, jns)).astype(np.float32)
    slpc = rng.uniform(98000, 104000, (iew, jns)).astype(np.float32)
    slpx = (slpc + rng.normal(0, 300, (iew, jns))).astype(np.float32)
    np.testing.assert_array_equal(oracle.pres_sfc_from_slp(pres, slpc, slpx), ctx.pres_sfc_from_slp(pres, slpc, slpx))


def _epilogue_case(kbu_levels, iew, jns, seed):
    rng = np.random.default_rng(seed)
    press = np.array(kbu_levels, np.float32)
    kbu = len(press)
    t = np.empty((kbu, iew, jns), np.float32); h = np.empty_like(t)
    for k, p in enumerate(press):
        t[k] = 288.0 * (min(p, 1000.0) / 1000.0) ** 0.19 + rng.normal(0, 1.5, (iew, jns))
        h[k] = 44330.0 * (1 - (min(p, 1000.0) / 1013.25) ** 0.1903) + rng.normal(0, 10, (iew, jns))
    rh = rng.uniform(-5, 110, (kbu, iew, jns)).astype(np.float32)
    terrain = rng.uniform(0, 5200, (iew, jns)).astype(np.float32)
    slp = rng.uniform(99000, 103000, (iew, jns)).astype(np.float32)
    return press, rh, t, h, terrain, slp


@pytest.mark.parametrize("iew,jns", [(9, 8), (150, 131)])
def test_mixing_ratio_and_sfcprs(ctx, oracle, iew, jns):
    """final_analysis_module.F90:1026-1201.  EXP / LOG / ** are glibc's expf / logf / powf in the
    reference build and FP64-then-round on the device: tolerance 1e-6 relative (a few FP32 ulp
    through the chain of powers), and the clamped rh is bit-identical."""
    press, rh, t, h, terrain, slp = _epilogue_case([1001, 1000, 925, 850, 700, 500, 400, 300], iew, jns, iew)
    rc, rh0, qv0, ps0 = oracle.mixing_ratio(rh, t, h, terrain, slp, press)
    assert rc == 0
    rh1, qv1, ps1 = ctx.mixing_ratio(rh, t, h, terrain, slp, press)
    np.testing.assert_array_equal(rh0, rh1)
    np.testing.assert_allclose(ps1, ps0, rtol=1e-6, atol=0)
    np.testing.assert_allclose(qv1, qv0, rtol=1e-6, atol=1e-12)
    # almost every value is in fact identical: the differences are isolated last-bit cases
    assert np.mean(ps1 != ps0) < 0.02 and np.mean(qv1 != qv0) < 0.02


def test_mixing_ratio_needs_the_mandatory_levels(ctx):
    from obsgrid_b200._capi import OgError
    press, rh, t, h, terrain, slp = _epilogue_case([1001, 1000, 925, 850], 6, 6, 1)
    with pytest.raises(OgError):
        ctx.mixing_ratio(rh, t, h, terrain, slp, press)


def test_qc_time_fdda_density_bit_exact(ctx, oracle):
    """proc_qc of a surface-FDDA time (proc_qc.F90:243-420): flags as og_qc_time, tobbox and odis
    bit-identical (counts are exact in FP32; the distance is an exact EDT)."""
    base = synth.small_case(iew=90, jns=70, kbu=5, n_sfc=60, n_snd=10, radii=(6, 4), seed=23,
                            dx_m=60000.0)      # 250 km = 4.2 cells: boxes with and without obs
    rng = np.random.default_rng(2)
    tb0 = rng.integers(0, 3, (base.iew, base.jns)).astype(np.float32)   # the driver hands in zeros (driver.F90:140); any start must work
    for tb_in in (np.zeros_like(tb0), tb0):
        cpu, gpu = base.copy(), base.copy()
        tb_c, od_c = oracle.qc_time_fdda(cpu, tb_in)
        tb_g, od_g = ctx.proc_qc_fdda(gpu, tb_in)
        np.testing.assert_array_equal(tb_c, tb_g)
        np.testing.assert_array_equal(od_c, od_g)
        assert tb_c.max() >= 1.0 and od_c.max() > 0.0 and (tb_c == 0).any()
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
            np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm))


class _OracleProc:
    """The oracle behind the interface fdda.HostBackend drives."""

    def __init__(self, oracle):
        self.o = oracle

    def proc_qc_fdda(self, case, tob):
        return self.o.qc_time_fdda(case, tob)

    def proc_oa(self, case):
        return self.o.oa_time(case)

    def temporal_interp(self, a, b, cf, c1, c2, cross):
        rc, out = self.o.temporal_interp(a, b, cf, c1, c2, cross)
        assert rc == 0
        return out


def test_surface_fdda_chain(ctx, oracle):
    """BASELINE configs[3] as the pipeline it is (driver.F90:107-113, :306-341, :561-563): 3-D analyses of
    the traditional times -> store_fa -> temporal_interp -> proc_qc with the observation density ->
    surface-only proc_oa for every FDDA time in between, tobbox / odis cleared at every time (driver.F90:140-141).
      * the GPU chain with host arrays and the GPU chain with everything resident (fdda.DeviceBackend)
        give the same bits;
      * against the oracle composition: every flag, t and slp bit for bit along the whole chain;
      * u / v / rh within the north-star tolerance stage by stage -- each FDDA time re-analysed by the
        GPU from the oracle's own interpolated first guess (a tolerance-level difference of a wind field
        that is then interpolated, checked and re-analysed can flip a circle / ellipse / banana
        decision further down the chain, so the end-to-end comparison of those fields is not a
        per-stage bound)."""
    import torch
    from obsgrid_b200 import fdda
    small = dict(iew=72, jns=60, kbu=5, n_sfc=500, n_snd=60, radii=(6, 4), dx_m=60000.0)
    ratio = 3

    class Recording(fdda.HostBackend):
        """Keeps every FDDA time's input state (first guess, flags, tobbox) of the oracle chain."""
        def __init__(self, proc):
            super().__init__(proc); self.inputs = []
        def analyse_surface(self, hc, tob):
            self.inputs.append((hc.copy(), tob.copy()))
            return super().analyse_surface(hc, tob)

    runs = {}
    for name in ("oracle", "host", "device"):
        cases3d, hourly = fdda.synthetic_chain(None, 3, ratio, **small)
        if name == "device":
            be = fdda.DeviceBackend(ctx, torch, torch.device("cuda", 0))
            for c in cases3d + hourly:
                be.upload(c)
        elif name == "oracle":
            be = rec = Recording(_OracleProc(oracle))
        else:
            be = fdda.HostBackend(ctx)
        fdda.run_chain(be, cases3d, hourly, ratio)
        if name == "device":
            ctx.sync()
            for c in cases3d + hourly:
                be.download(c)
        runs[name] = (cases3d, hourly)
    assert len(runs["oracle"][1]) == 2 * (ratio - 1)
    for a, b in zip(runs["host"][0] + runs["host"][1], runs["device"][0] + runs["device"][1]):
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
            np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=f"host vs device:{nm}")
    for a, b in zip(runs["oracle"][0] + runs["oracle"][1], runs["host"][0] + runs["host"][1]):
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "slp"):
            np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=f"oracle vs gpu:{nm}")
    for a, b in zip(runs["oracle"][0], runs["host"][0]):            # the 3-D times: same inputs on both sides
        for nm in ("u", "v", "rh"):
            np.testing.assert_allclose(getattr(b, nm), getattr(a, nm), rtol=RTOL, atol=ATOL, err_msg=nm)
    for (inp, tob), want in zip(rec.inputs, runs["oracle"][1]):     # the FDDA times: stage by stage
        got = inp.copy()
        tb, _ = ctx.proc_qc_fdda(got, tob)
        ctx.proc_oa(got)
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "slp"):
            np.testing.assert_array_equal(getattr(got, nm), getattr(want, nm), err_msg=nm)
        for nm in ("u", "v", "rh"):
            np.testing.assert_allclose(getattr(got, nm), getattr(want, nm), rtol=RTOL, atol=ATOL, err_msg=nm)
    h0 = runs["oracle"][1][0]
    fresh = fdda.synthetic_chain(None, 3, ratio, **small)[1][0]
    assert np.any(h0.t != fresh.t) and np.count_nonzero(h0.qc_t) > 0


@pytest.mark.parametrize("shape", [(3, 6, 5), (27, 120, 97), (64, 51)])
def test_temporal_interp_bit_exact(ctx, oracle, shape):
    """tinterp_module.F90:53-215, every weight pattern of a 6-hour bracket, both grid kinds."""
    rng = np.random.default_rng(len(shape) * 17 + shape[-1])
    a = rng.normal(280, 5, shape).astype(np.float32); b = rng.normal(282, 5, shape).astype(np.float32)
    for (cf, c1, c2) in ((1, 0, 6), (2, 0, 6), (5, 0, 6), (0, 0, 6), (6, 0, 6), (3, 3, 3)):
        for cross in (False, True):
            rc, want = oracle.temporal_interp(a, b, cf, c1, c2, cross)
            assert rc == 0
            np.testing.assert_array_equal(ctx.temporal_interp(a, b, cf, c1, c2, cross), want)
    from obsgrid_b200._capi import OgError
    with pytest.raises(OgError):
        ctx.temporal_interp(a, b, 4, 2, 2, False)


# ---- OG_ARITH_CONTRACTED: the opt-in scan arithmetic (FMA contraction, hardware reciprocal) --------
@pytest.fixture
def contracted(ctx):
    ctx.set_arithmetic("contracted")
    assert ctx.arithmetic() == "contracted"
    yield ctx
    ctx.set_arithmetic("exact")


@pytest.mark.parametrize("var,pressure", [(3, 500.0), (5, 1001.0), (1, 300.0), (2, 850.0), (4, 1001.0)])
def test_contracted_cressman_within_north_star_tolerance(contracted, oracle, var, pressure):
    """Every shape (circles for TT/PMSL; circles, ellipses and bananas for UU/VV/RH, with RH
    scaling) against the oracle at 1e-5 relative / 1e-3 absolute, the north-star bar; the
    observed deviation is two orders of magnitude smaller, which the test also pins."""
    rng = np.random.default_rng(int(pressure) * 7 + var)
    iew, jns, n, R = 140, 110, 2500, 7
    g = wavy(iew, jns, 20.0, 50.0)
    ub, vb = wind_fields(iew, jns)
    x, y = rng_obs(rng, n, iew, jns, lo=1.0)
    d = rng.normal(0, 5, n).astype(np.float32)
    fg = np.abs(rng.normal(50, 20, n)).astype(np.float32) + 1
    a = oracle.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    b = contracted.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    np.testing.assert_allclose(b, a, rtol=RTOL, atol=ATOL)
    assert np.max(np.abs(b.astype(np.float64) - a)) < 1e-4
    assert np.any(a != g)


def test_contracted_time_small(contracted, oracle):
    """A whole analysis time: flags are untouched by the mode (QC is always exact); all five
    analysed fields within the north-star tolerance of the oracle."""
    base = synth.small_case(iew=60, jns=52, kbu=6, n_sfc=500, n_snd=120, radii=(6, 4, 2), scale_rh=True)
    cpu, gpu = base.copy(), base.copy()
    oracle.qc_time(cpu); contracted.proc_qc(gpu)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm), err_msg=nm)
    assert oracle.oa_time(cpu) == contracted.proc_oa(gpu) == 0
    for nm in ("t", "u", "v", "rh", "slp"):
        np.testing.assert_allclose(getattr(gpu, nm), getattr(cpu, nm), rtol=RTOL, atol=ATOL, err_msg=nm)


def test_arithmetic_mode_is_validated(ctx):
    from obsgrid_b200._capi import lib
    assert lib().og_set_arithmetic(ctx._h, 7) == -2 and ctx.arithmetic() == "exact"


def test_time_slab_sharding_is_exact(ctx):
    """Level x variable sharding (obsgrid_b200/shard.py, north star: 'partitioned by pressure level x
    variable x analysis time'): the per-rank programs of run_time_sharded_slabs, run one after the
    other on this GPU and merged as the collectives would (OR of the surface flag rows, gather),
    equal the unsharded time bit for bit."""
    from obsgrid_b200 import shard
    base = synth.small_case(iew=50, jns=44, kbu=7, n_sfc=300, n_snd=100, radii=(5, 3))
    full = base.copy()
    ctx.proc_qc(full); ctx.proc_oa(full)

    def qc(c, k0, k1, cons, mask):
        c.qc_params.var_mask = mask; ctx.proc_qc(c, k0, k1, cons); c.qc_params.var_mask = 0

    def oa(c, k0, k1, mask):
        c.oa_params.var_mask = mask; ctx.proc_oa(c, k0, k1); c.oa_params.var_mask = 0

    for world in (2, 4, 7):        # 7: RH and DEWPOINT of the surface on different ranks
        plan = shard.plan_slabs(base, world)
        cs = [base.copy() for _ in range(world)]
        for c, p in zip(cs, plan):
            if p["sfc_mask"]:
                qc(c, 0, 1, False, p["sfc_mask"])
            if p["k1"] > p["k0"]:
                qc(c, p["k0"], p["k1"], True, 0)
        rows = [np.bitwise_or.reduce([c.qc_slp if nm == "qc_slp" else getattr(c, nm)[0] for c in cs], axis=0)
                for nm in shard.SFC_FLAGS]
        out = base.copy()
        for c, p in zip(cs, plan):
            c.qc_u[0], c.qc_v[0], c.qc_t[0], c.qc_rh[0] = rows[0], rows[1], rows[2], rows[3]
            c.qc_slp[...] = rows[4]
            ctx.qc_consistency(c, 0, 1)
            if p["sfc_mask"] & 0x1f:
                oa(c, 0, 1, p["sfc_mask"] & 0x1f)
            if p["k1"] > p["k0"]:
                oa(c, p["k0"], p["k1"], 0)
            for nm in ("t", "u", "v", "rh", "qc_u", "qc_v", "qc_t", "qc_rh"):
                getattr(out, nm)[p["k0"]:p["k1"]] = getattr(c, nm)[p["k0"]:p["k1"]]
            for nm, bits in shard.SFC_VARS:
                if p["sfc_mask"] & bits & 0x1f:
                    if nm == "slp":
                        out.slp[...] = c.slp
                    else:
                        getattr(out, nm)[0] = getattr(c, nm)[0]
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh"):
            getattr(out, nm)[0] = getattr(cs[0], nm)[0]
        out.qc_slp[...] = cs[0].qc_slp
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
            np.testing.assert_array_equal(getattr(full, nm), getattr(out, nm), err_msg=f"{nm} world {world}")


def test_dev_item_lists_equal_the_range_calls(ctx):
    """og_dev_qc_time_items / og_dev_oa_time_items (one batch over "surface slabs + a run of upper
    levels") == the corresponding range calls with var_mask, bit for bit."""
    import torch
    base = synth.small_case(iew=50, jns=44, kbu=7, n_sfc=300, n_snd=100, radii=(5, 3))
    sfc_mask = (1 << 0) | (1 << 3) | (1 << 6)            # UU, RH (+ DEWPOINT) of the surface level
    ref = base.copy()
    ref.qc_params.var_mask = sfc_mask; ctx.proc_qc(ref, 0, 1, False); ref.qc_params.var_mask = 0
    ctx.proc_qc(ref, 3, 6, True)
    ctx.qc_consistency(ref, 0, 1)
    ref.oa_params.var_mask = sfc_mask & 0x1f; ctx.proc_oa(ref, 0, 1); ref.oa_params.var_mask = 0
    ctx.proc_oa(ref, 3, 6)

    dev = torch.device("cuda", 0)
    names = ("t", "u", "v", "rh", "slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp",
             "xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")
    wk = {nm: torch.from_numpy(getattr(base, nm).copy()).to(dev) for nm in names}
    d = base.desc()
    for nm in names:
        setattr(d, nm, wk[nm].data_ptr())
    levels, masks = [0, 3, 4, 5], [sfc_mask, 0, 0, 0]
    ctx.dev_proc_qc_items(d, base.qc_params, levels, masks, True)
    ctx.dev_qc_consistency(d, 0, 1)
    ctx.dev_proc_oa_items(d, base.oa_params, levels, [m & 0x1f for m in masks])
    ctx.sync()
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(wk[nm].cpu().numpy(), getattr(ref, nm), err_msg=nm)
    assert np.any(ref.u[0] != base.u[0]) and np.array_equal(ref.v[0], base.v[0]) and np.array_equal(ref.t[1], base.t[1])
    # malformed lists are reported, not executed
    from obsgrid_b200._capi import OgError
    with pytest.raises(OgError):
        ctx.dev_proc_qc_items(d, base.qc_params, [3, 2], [0, 0])


def test_no_qc_test_selected_leaves_only_the_consistency_pass(ctx, oracle):
    """any_checks_to_do (proc_qc.F90:237): with error_max and buddy_check both off the level x variable
    loop is skipped -- no flags from the tests, no dew point stored, no observation density -- while
    tobbox / 5, obs_distance and qc_consistency still run."""
    from obsgrid_b200 import perslab
    base = synth.small_case(iew=50, jns=44, kbu=4, n_sfc=200, n_snd=40, radii=(5, 3), seed=31)
    base.qc_params.qc_test_error_max = 0
    base.qc_params.qc_test_buddy = 0
    base.qc_u[0, ::7] |= 1 << 17                     # something for qc_consistency to copy to V
    cpu, gpu, slab = base.copy(), base.copy(), base.copy()
    tb0 = np.zeros((base.iew, base.jns), np.float32)
    tb_c, od_c = oracle.qc_time_fdda(cpu, tb0)
    tb_g, od_g = ctx.proc_qc_fdda(gpu, tb0)
    perslab.proc_qc_per_slab(ctx, slab)
    assert not tb_c.any()
    np.testing.assert_array_equal(tb_c, tb_g)
    np.testing.assert_array_equal(od_c, od_g)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm), err_msg=nm)
        np.testing.assert_array_equal(getattr(cpu, nm), getattr(slab, nm), err_msg="per slab " + nm)
    assert np.array_equal(gpu.qc_t, base.qc_t) and np.any(gpu.qc_v != base.qc_v)


def test_upper_levels_under_surface_only_are_a_no_op(ctx):
    """proc_oa.F90:365: with fdda_loop > 1 only the surface level is analysed.  A batch that holds
    upper levels only then has no slab at all -- it must come back clean, fields untouched."""
    base = synth.small_case(iew=40, jns=36, kbu=4, n_sfc=120, n_snd=30, radii=(5, 3))
    base.oa_params.surface_only = 1
    got = base.copy()
    ctx.proc_oa(got, 1, 4)
    for nm in ("t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(getattr(got, nm), getattr(base, nm), err_msg=nm)
    ctx.proc_oa(got, 0, 4)
    assert np.any(got.t[0] != base.t[0]) and np.array_equal(got.t[1:], base.t[1:])


def test_context_on_another_device_from_a_fresh_thread(oracle):
    """A host thread that never touched CUDA has device 0 current; a context on another GPU must
    still work from it (every entry point binds the context's device).  Needs two GPUs."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from obsgrid_b200.api import Context
    c1 = Context(1)
    base = synth.small_case(iew=40, jns=36, kbu=4, n_sfc=120, n_snd=30, radii=(5, 3))
    got, err = base.copy(), []

    def work():
        try:
            c1.proc_qc_oa(got)
            c1.time_upload(got, 0); c1.time_compute(got, 0); c1.time_download(got, 0); c1.time_wait(0)
            c1.sync()
        except BaseException as e:      # noqa: BLE001
            err.append(e)

    t = threading.Thread(target=work); t.start(); t.join()
    c1.close()
    assert not err, err
    # the asynchronous pass re-analysed the output of the blocking one: flags are those of the
    # first QC (a second QC of unchanged obs sets the same bits), fields moved twice
    ref = base.copy()
    oracle.qc_time(ref)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_slp"):
        assert np.all((getattr(got, nm) & getattr(ref, nm)) == getattr(ref, nm)), nm
    assert np.any(got.t != base.t)


class _ThreadRanks:
    """`world` contexts on one GPU standing in for `world` ranks: the exchange callback combines the
    buffers of the threads through host memory (barrier, reduce, barrier)."""

    def __init__(self, world):
        import threading
        import torch
        from obsgrid_b200 import _capi
        self.torch, self.capi, self.world = torch, _capi, world
        self.barrier = threading.Barrier(world)
        self.stage = [None] * world
        self.calls = [0] * world

    def callback(self, rank):
        torch, capi = self.torch, self.capi
        from obsgrid_b200.shard import _RawBuffer

        def fn(_user, buf, count, dtype, op, stream):
            try:
                ts = "|u1" if dtype == capi.OG_XCHG_UINT8 else "<i4"
                t = torch.as_tensor(_RawBuffer(buf, count, ts), device="cuda:0")
                s = torch.cuda.ExternalStream(stream, device="cuda:0")
                with torch.cuda.stream(s):
                    self.stage[rank] = t.cpu()
                s.synchronize()
                self.barrier.wait(timeout=60)
                stack = torch.stack(self.stage)
                red = stack.max(dim=0).values if op == capi.OG_XCHG_MAX else stack.sum(dim=0).to(t.dtype)
                self.barrier.wait(timeout=60)
                with torch.cuda.stream(s):
                    t.copy_(red.to("cuda:0"))
                s.synchronize()
                self.calls[rank] += 1
                return 0
            except BaseException as e:   # noqa: BLE001
                print("exchange failed", repr(e))
                self.barrier.abort()
                return 1
        return capi.OG_EXCHANGE_FN(fn)

    def run(self, work):
        """work(ctx, rank) -> result, one thread per rank; returns the results."""
        import threading
        from obsgrid_b200.api import Context
        out, err = [None] * self.world, []

        def body(r):
            try:
                c = Context(0)
                c.set_exchange(r, self.world, self.callback(r))
                out[r] = work(c, r)
                c.sync()
                c.close()
            except BaseException as e:   # noqa: BLE001
                err.append(e)
                self.barrier.abort()
        th = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        [t.start() for t in th]; [t.join() for t in th]
        assert not err, err
        return out


@pytest.mark.parametrize("world", [2, 3])
def test_buddy_cells_dealt_to_ranks_equal_one_gpu(ctx, oracle, world):
    """og_set_exchange: the buddy check of ONE slab with its cells dealt to `world` ranks (threads on
    one GPU here; tests/test_shard_gpu.py runs real ranks over NCCL).  A sparse network that needs
    several passes of the relaxation fix-point: every rank must end with the flags of the
    single-GPU run -- which are the oracle's."""
    rng = np.random.default_rng(8)
    iew, jns, n = 400, 400, 1500
    g = np.zeros((jns, iew), np.float32)
    x, y = rng_obs(rng, n, iew, jns)
    ob = rng.normal(0, 6, n).astype(np.float32)
    ob[rng.random(n) < 0.3] += 11.0
    z = np.zeros(n, np.float32)
    q0 = (rng.integers(0, 2, n) * (1 << 13)).astype(np.int32)
    args = (ob, x, y, q0, z, g, 5, 500.0, 30.0, 1.0, 8.0)
    a = oracle.buddy_check(*args)
    one = ctx.buddy_check(*args)
    np.testing.assert_array_equal(a[0], one[0])
    ranks = _ThreadRanks(world)
    got = ranks.run(lambda c, r: (c.buddy_check(*args), c.stats()))
    for r in range(world):
        np.testing.assert_array_equal(got[r][0][0], a[0], err_msg=f"rank {r}")
        assert got[r][1]["buddy_fixpoint_iters"] >= 2
    # the pair work was dealt out, not replicated
    pe = [got[r][1]["buddy_pairs_examined"] for r in range(world)]
    assert sum(pe) == ctx_pairs(ctx, args) and max(pe) < sum(pe)
    assert all(c >= 3 for c in ranks.calls)        # a mask exchange per pass + the flags


def ctx_pairs(ctx, args):
    ctx.reset_stats()
    ctx.buddy_check(*args)
    return ctx.stats()["buddy_pairs_examined"]


@pytest.mark.parametrize("world,smooth", [(2, 0), (3, 0), (2, 2)])
def test_time_with_cells_dealt_to_ranks(ctx, world, smooth):
    """The batched time path under og_set_exchange: every (level, variable) slab is cut over the ranks
    -- the buddy check's cells (relaxation mask and flags combined by MAX) and the Cressman cells (the
    slab summed as int32 after every scan, so that every rank interpolates the next innovations from
    the complete field); with smoothing the perturbation is what travels.  Flags and analysed fields
    on every rank equal the single-GPU run bit for bit."""
    base = synth.small_case(iew=70, jns=50, kbu=5, n_sfc=900, n_snd=150, seed=4, radii=(6, 3))
    if smooth:
        p = base.oa_params
        p.smooth_type = 1
        p.smooth_sfc_wind = p.smooth_sfc_temp = p.smooth_sfc_rh = p.smooth_sfc_slp = smooth
        p.smooth_upper_wind = p.smooth_upper_temp = p.smooth_upper_rh = 1
    ref = base.copy()
    ctx.proc_qc_oa(ref)

    def work(c, r):
        got = base.copy()
        c.proc_qc_oa(got)
        return got
    ranks = _ThreadRanks(world)
    for r, got in enumerate(ranks.run(work)):
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
            np.testing.assert_array_equal(getattr(got, nm), getattr(ref, nm), err_msg=f"rank {r} {nm}")
    assert np.any(ref.t != base.t) and np.any(ref.u != base.u)
    assert all(c > 4 for c in ranks.calls)
