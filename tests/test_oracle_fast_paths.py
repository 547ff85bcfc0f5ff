"""The oracle's fast paths for the large configurations (oracle/og_oracle.c: ogo_set_fast_paths)
re-schedule the as-written loops -- spatial bins instead of the N^2 station loops of buddy_check
(qc2_module.F90:223-329), grid rows / target bins dealt to pthreads -- and must reproduce them bit
for bit: they are what checks the GPU at 200 k - 2 M reports, where the N^2 loops need hours."""
import numpy as np
import pytest

from obsgrid_b200 import synth


@pytest.fixture
def fast(oracle):
    yield oracle
    oracle.set_fast_paths(False, 1)


def _buddy_case(rng, iew, jns, n, clustered, spread):
    x = rng.uniform(2, iew - 1, n); y = rng.uniform(2, jns - 1, n)
    if clustered:
        k = n // 3
        x[:k] = np.clip(rng.normal(iew * 0.4, iew * 0.05, k), 2, iew - 1)
        y[:k] = np.clip(rng.normal(jns * 0.6, jns * 0.05, k), 2, jns - 1)
        p = rng.permutation(n); x, y = x[p], y[p]
    x = x.astype(np.float32); y = y.astype(np.float32)
    if n > 10:       # duplicates (distance 0 < 0.1 km: not buddies) and near-duplicates
        x[5], y[5] = x[3], y[3]
        x[9], y[9] = x[3] + np.float32(1e-3), y[3]
    jj, ii = np.meshgrid(np.arange(jns), np.arange(iew), indexing="ij")
    g = (280 + 6 * np.sin(ii / 17.0) * np.cos(jj / 13.0)).astype(np.float32)
    ob = (280 + rng.normal(0, spread, n)).astype(np.float32)
    ob[rng.random(n) < 0.05] += 25.0
    ob[rng.random(n) < 0.03] -= 25.0
    elev = rng.uniform(-500, 3000, n).astype(np.float32)
    return g, x, y, ob, elev


@pytest.mark.parametrize("iew,jns,n,var,pressure,dx,clustered,spread,threads", [
    (300, 260, 6000, 3, 850.0, 30.0, True, 5.0, 1),      # 300 km = 10 cells
    (300, 260, 6000, 3, 1001.0, 12.0, False, 5.0, 4),    # 500 km = 42 cells, threads
    (120, 100, 3000, 6, 1001.0, 9.0, True, 300.0, 3),    # PMSLPSFC: difmax depends on the elevation (negative too)
    (400, 400, 1500, 5, 500.0, 30.0, False, 6.0, 2),     # sparse: many stations keep exactly two buddies
    (40, 40, 5000, 7, 150.0, 50.0, True, 8.0, 4),        # range wider than the grid: one bin
    (64, 64, 1, 1, 700.0, 30.0, False, 5.0, 2),
    (64, 64, 0, 1, 700.0, 30.0, False, 5.0, 2),
])
def test_binned_buddy_check_equals_the_station_loops(fast, iew, jns, n, var, pressure, dx, clustered, spread, threads):
    rng = np.random.default_rng(n + var)
    g, x, y, ob, elev = _buddy_case(rng, iew, jns, n, clustered, spread)
    if var == 5 and n > 100:
        ob[rng.random(n) < 0.3] += 11.0       # diffs between 1.25*difmax and 1.25*1.2*difmax
    if var == 7:
        ob = (ob - 45).astype(np.float32)       # every dew-point factor band
    q0 = rng.integers(0, 2, n).astype(np.int32) * 4
    fast.set_fast_paths(False, 1)
    fast.reset_stats()
    a = fast.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, 1.0, 8.0)
    full = fast.get_stats()
    fast.set_fast_paths(True, threads)
    fast.reset_stats()
    b = fast.buddy_check(ob, x, y, q0, elev, g, var, pressure, dx, 1.0, 8.0)
    st = fast.get_stats()
    for nm, u, v in zip(("qc", "aob", "err", "difmax", "buddy_num_final"), a, b):
        np.testing.assert_array_equal(u, v, err_msg=nm)
    assert st["buddy_pair_tests"] == full["buddy_pair_tests"] == n * max(n - 1, 0)
    if n >= 3000 and iew >= 300:
        assert 0 < st["buddy_pairs_examined"] < full["buddy_pair_tests"]
    if var == 5 and n > 100:
        assert np.count_nonzero(a[4] == 2) > 10       # the relaxation path is exercised


@pytest.mark.parametrize("var,pressure,threads", [(3, 500.0, 4), (1, 300.0, 3), (4, 1001.0, 5), (2, 850.0, 8)])
def test_row_banded_cressman_equals_the_single_loop(fast, var, pressure, threads):
    rng = np.random.default_rng(var * 10 + threads)
    iew, jns, n, R = 150, 131, 3000, 7
    jj, ii = np.meshgrid(np.arange(jns), np.arange(iew), indexing="ij")
    g = (50 + 20 * np.sin(ii / 11.0) * np.cos(jj / 9.0)).astype(np.float32)
    ub = (30 * np.sin(jj / 20.0) + 5).astype(np.float32); vb = (25 * np.cos(ii / 25.0)).astype(np.float32)
    x = rng.uniform(1, iew, n).astype(np.float32); y = rng.uniform(1, jns, n).astype(np.float32)
    d = rng.normal(0, 5, n).astype(np.float32)
    fg = (np.abs(rng.normal(50, 20, n)) + 1).astype(np.float32)
    fast.set_fast_paths(False, 1); fast.reset_stats()
    a = fast.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    sa = fast.get_stats()
    fast.set_fast_paths(False, threads); fast.reset_stats()
    b = fast.cressman(d, x, y, g, var, R, 12.0, ub, vb, pressure, fg_value=fg, scale_rh=True)
    sb = fast.get_stats()
    np.testing.assert_array_equal(a, b)
    assert sa == sb and sa["cressman_updates"] > 0
    if var != 3:
        assert sa["n_ellipse"] + sa["n_banana"] > 0


def test_whole_time_with_fast_paths_equals_the_as_written_oracle(fast):
    base = synth.small_case(iew=90, jns=70, kbu=5, n_sfc=900, n_snd=200, radii=(6, 4, 2), scale_rh=True)
    a, b = base.copy(), base.copy()
    fast.set_fast_paths(False, 1)
    fast.qc_time(a); fast.oa_time(a)
    fast.set_fast_paths(True, 4)
    fast.qc_time(b, threads=4); fast.oa_time(b, threads=4)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp"):
        np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)
    assert np.count_nonzero(a.qc_t) > 0
