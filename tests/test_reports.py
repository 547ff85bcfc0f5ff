"""Report list <-> flat observation table (include/obsgrid_reports.h).

The reference asks proc_ob_access for every (level, variable) slab and walks every report's
measurement list each time (proc_ob_access.F90:146-403, ob_access_module.F90:11-817).  The table
builder decides the slab-independent part once per analysis time.  These tests hold it to the
oracle's restatement of the per-slab routines as written (oracle/og_oracle_reports.c):

  * every slab's 'get' and 'use' return exactly what the device-side selection rule makes of the
    table rows (same observations, same order, same values and flags);
  * a whole proc_qc over the report list, as written, leaves the same flags (and dew points, discard
    marks, extend_influence / no_qc_possible side effects) in the list as table -> QC -> put.
"""
import numpy as np
import pytest

from obsgrid_b200 import reports as R
from obsgrid_b200 import synth

MISSING = R.MISSING
FAILS_ERROR_MAX, FAILS_BUDDY = 1 << 16, 1 << 17


def _case(seed=11, **kw):
    args = dict(iew=70, jns=60, kbu=8, n_sfc=700, n_snd=260, radii=(6, 4), seed=seed)
    args.update(kw)
    return synth.small_case(**args)


def _not_missing(a):
    return np.abs(a - np.float32(MISSING)) > 1.0


def _select_from_table(t, k, var, qc_max, is_get):
    """The rule of og_time.cu:sel_pred on the table rows of level k."""
    ob = {1: t.ob_u, 2: t.ob_v, 3: t.ob_t, 4: t.ob_rh, 7: t.ob_rh}
    qc = {1: t.qc_u, 2: t.qc_v, 3: t.qc_t, 4: t.qc_rh, 7: t.qc_rh}
    if var == 5:
        v, q = t.ob_slp, t.qc_slp
    else:
        v, q = ob[var][k], qc[var][k]
    ok = _not_missing(v) & (q < qc_max)
    if var == 7:
        ok &= _not_missing(t.ob_t[k]) & (t.qc_t[k] < qc_max)
    ok &= (t.xob >= 2) & (t.xob <= t.iew - 1) & (t.yob >= 2) & (t.yob <= t.jns - 1)
    if is_get:
        ok &= t.bogus == 0
    return np.flatnonzero(ok), v, q


@pytest.mark.parametrize("p_diff", [1, 1500])
def test_every_slab_selection_equals_proc_ob_access(oracle, p_diff):
    case = _case()
    rl, date, time = R.synthetic_reports(case, seed=3)
    t = R.reports_to_table(rl.copy(), case.iew, case.jns, case.pressure, date, time, p_diff)
    nsel = 0
    for is_get, qc_max in ((True, 1 << 20), (False, FAILS_ERROR_MAX)):
        for k in range(case.kbu):
            for var in (1, 2, 3, 4, 5, 7):
                if var == 5 and k > 0:
                    continue
                if var == 7 and not is_get:
                    continue
                got = oracle.get_obs(rl.copy(), is_get, var, float(case.pressure[k]), date, time, p_diff, qc_max,
                                     case.iew, case.jns)
                idx, v, q = _select_from_table(t, k, var, qc_max, is_get)
                np.testing.assert_array_equal(got["index"], idx, err_msg=f"get={is_get} k={k} var={var}")
                if var == 7:
                    want = np.array([oracle.dewpoint(a, b) for a, b in zip(t.ob_t[k][idx], t.ob_rh[k][idx])], np.float32)
                else:
                    want = v[idx]
                np.testing.assert_array_equal(got["value"], want)
                np.testing.assert_array_equal(got["qc"], q[idx])
                np.testing.assert_array_equal(got["x"], t.xob[idx])
                np.testing.assert_array_equal(got["elev"], t.elev[idx])
                nsel += idx.size
    assert nsel > 2000


def test_builder_side_effects_match_query_ob(oracle):
    """no_qc_possible on reports outside the time window, extend_influence on single-level reports
    matched within the tolerance: the list after the builder == the list after the as-written queries."""
    case = _case(seed=5)
    rl, date, time = R.synthetic_reports(case, seed=8)
    for r in rl.reports:
        r.bogus = 0          # (bogus reports are first asked by proc_oa: covered by the whole-run test)
    a, b = rl.copy(), rl.copy()
    R.reports_to_table(a, case.iew, case.jns, case.pressure, date, time, 1500)
    for is_get, qc_max, variables in ((True, 1 << 20, (1, 2, 3, 4, 5, 7)), (False, FAILS_ERROR_MAX, (1, 2, 3, 4, 5))):
        for k in range(case.kbu):            # proc_qc's requests, then proc_oa's, in the reference's order
            for var in variables:
                if var == 5 and k > 0:
                    continue
                oracle.get_obs(b, is_get, var, float(case.pressure[k]), date, time, 1500, qc_max, case.iew, case.jns)
    np.testing.assert_array_equal(a.meas, b.meas)
    assert np.count_nonzero(a.meas["pressure"]["qc"] & 4) > 0 and np.count_nonzero(a.meas["u"]["qc"] == (1 << 15)) > 0
    assert np.any(a.meas != rl.meas)


def _qc_through_table(case, rl, date, time, p_diff, run_qc):
    t = R.reports_to_table(rl, case.iew, case.jns, case.pressure, date, time, p_diff)
    c = case.copy()
    for n in R.Table.NAMES:
        setattr(c, n, getattr(t, n))
    run_qc(c)
    R.table_to_reports(t, rl, date, time)
    return t


def _use_requests_of_proc_oa(oracle, case, rl, date, time, p_diff):
    """The 'use' requests proc_oa makes after proc_qc (proc_oa.F90:455-470): they are the first to
    meet bogus reports ('get' skips those), with the side effects of any query_ob call."""
    for k in range(case.kbu):
        for var in (1, 2, 3, 4, 5):
            if var == 5 and k > 0:
                continue
            oracle.get_obs(rl, False, var, float(case.pressure[k]), date, time, p_diff, FAILS_ERROR_MAX, case.iew, case.jns)


def _assert_lists_equal(a, b):
    for fld in ("u", "v", "temperature", "rh", "dew_point", "speed", "direction", "pressure", "height"):
        np.testing.assert_array_equal(a.meas[fld]["qc"], b.meas[fld]["qc"], err_msg=f"{fld} qc")
        np.testing.assert_array_equal(a.meas[fld]["data"], b.meas[fld]["data"], err_msg=f"{fld} data")
    np.testing.assert_array_equal(a.header("discard"), b.header("discard"))
    np.testing.assert_array_equal([r.slp.qc for r in a.reports], [r.slp.qc for r in b.reports])


@pytest.mark.parametrize("p_diff,seed", [(1, 2), (1500, 4)])
def test_proc_qc_over_the_list_equals_table_qc_put(oracle, p_diff, seed):
    case = _case(seed=seed)
    rl, date, time = R.synthetic_reports(case, seed=seed + 10)
    a, b = rl.copy(), rl.copy()
    assert oracle.qc_time_reports(case, a, date, time, p_diff) == 0
    _use_requests_of_proc_oa(oracle, case, a, date, time, p_diff)
    _qc_through_table(case, b, date, time, p_diff, lambda c: oracle.qc_time(c))
    _assert_lists_equal(a, b)
    # the run did something: flags of every kind, dew points stored, td > t rule, discards
    m = a.meas
    assert np.count_nonzero(m["temperature"]["qc"] & FAILS_ERROR_MAX) > 0
    assert np.count_nonzero(m["rh"]["qc"] & FAILS_BUDDY) > 0
    assert np.count_nonzero(m["dew_point"]["qc"]) > 0 and np.any(m["dew_point"]["data"] != rl.meas["dew_point"]["data"])
    assert a.header("discard").sum() > rl.header("discard").sum()
    assert np.count_nonzero(m["speed"]["qc"]) > 0


@pytest.mark.gpu
def test_gpu_qc_from_a_report_list(ctx, oracle):
    """The same round trip with the GPU doing proc_qc on the table (all rows of ABI 3 in use)."""
    case = _case(seed=9, n_sfc=1500, n_snd=500)
    rl, date, time = R.synthetic_reports(case, seed=21)
    a, b = rl.copy(), rl.copy()
    assert oracle.qc_time_reports(case, a, date, time, 1) == 0
    _use_requests_of_proc_oa(oracle, case, a, date, time, 1)
    _qc_through_table(case, b, date, time, 1, lambda c: ctx.proc_qc(c))
    _assert_lists_equal(a, b)


@pytest.mark.gpu
def test_gpu_qc_oa_pipelines_carry_the_optional_rows(ctx, oracle):
    """og_qc_oa_time and the slot API move ob_td / qc_td / bogus like the other rows."""
    case = _case(seed=13)
    rl, date, time = R.synthetic_reports(case, seed=5)
    t = R.reports_to_table(rl.copy(), case.iew, case.jns, case.pressure, date, time, 1)
    base = case.copy()
    for n in R.Table.NAMES:
        setattr(base, n, getattr(t, n).copy())
    ref = base.copy(); oracle.qc_time(ref); oracle.oa_time(ref)
    one = base.copy(); ctx.proc_qc_oa(one)
    slot = base.copy(); out = base.copy()
    ctx.time_upload(slot, 0); ctx.time_compute(slot, 0); ctx.time_download(out, 0); ctx.time_wait(0)
    for got in (one, out):
        for n in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_td", "qc_slp", "ob_td", "t", "slp"):
            np.testing.assert_array_equal(getattr(got, n), getattr(ref, n), err_msg=n)
    assert np.any(ref.qc_td != base.qc_td)


def _round_trip_both_ways(oracle, case, rl, date, time, p_diff):
    a, b = rl.copy(), rl.copy()
    assert oracle.qc_time_reports(case, a, date, time, p_diff) == 0
    _use_requests_of_proc_oa(oracle, case, a, date, time, p_diff)
    t = _qc_through_table(case, b, date, time, p_diff, lambda c: oracle.qc_time(c))
    return a, b, t


def test_the_store_ob_defect_is_detected_not_reproduced(oracle):
    """ob_access_module.F90:708-740 against :399-402: query_ob skips the surface measurement when it looks
    for a pressure level, store_ob does not.  A sounding whose surface pressure lies within 1 Pa of an
    analysis level it also reports gets that level's checked value stored into its SURFACE measurement
    by the reference (profiles/exp/fuzz_reports.py found the case).  The table path does not reproduce
    the overwrite: it differs from the as-written calls on exactly those reports, which
    reports.store_ob_defect_reports names, and on no other."""
    case = _case(seed=21, kbu=5)
    rl, date, time = R.synthetic_reports(case, seed=33)
    for fld in ("u", "v", "temperature", "rh"):
        rl.meas[fld]["qc"][...] = 0                          # keep the example to the one effect
    lvl = np.float32(case.pressure[1]) * np.float32(100.0)
    for rep in rl.reports:
        if rep.n_meas < 3 or rep.bogus or rep.discard:
            continue
        m0, m1 = rep.first_meas, rep.first_meas + 1
        if abs(float(rl.meas["height"]["data"][m0]) - rep.elevation) > 0.5:
            continue                                         # no surface measurement first
        rl.meas["pressure"]["data"][m0] = lvl - np.float32(0.98)   # the surface, a hair off the analysis level
        rl.meas["pressure"]["data"][m1] = lvl - np.float32(0.9)    # and the sounding's own level there
        rl.meas["height"]["data"][m1] = rep.elevation + 300.0
        for f, v in (("u", 11.0), ("v", 7.0), ("temperature", 280.0), ("rh", 60.0)):
            rl.meas[f]["data"][m1] = v
    hit = R.store_ob_defect_reports(rl, case.pressure)
    assert len(hit) > 5
    a, b, t = _round_trip_both_ways(oracle, case, rl, date, time, 1)
    # the reference overwrote surface measurements ...
    over = [i for i in hit if a.meas["u"]["data"][rl.reports[i].first_meas] == np.float32(11.0)]
    assert len(over) > 3
    # ... the table path kept them, and agrees with the as-written calls everywhere else
    for i in hit:
        r = rl.reports[i]
        assert b.meas["u"]["data"][r.first_meas] == rl.meas["u"]["data"][r.first_meas]
    keep = np.ones(rl.meas.size, bool)
    for i in hit:
        r = rl.reports[i]
        keep[r.first_meas:r.first_meas + r.n_meas] = False
    for fld in ("u", "v", "temperature", "rh", "dew_point", "speed", "direction", "pressure", "height"):
        np.testing.assert_array_equal(a.meas[fld]["qc"][keep], b.meas[fld]["qc"][keep], err_msg=fld)
        np.testing.assert_array_equal(a.meas[fld]["data"][keep], b.meas[fld]["data"][keep], err_msg=fld)
    # without the coincidence there is no such report
    assert R.store_ob_defect_reports(R.synthetic_reports(_case(seed=5), seed=8)[0], _case(seed=5).pressure) == []


def test_one_measurement_answering_two_levels_is_reported(oracle):
    """use_p_tolerance_one_lev with a tolerance wider than half the spacing of two analysis levels:
    og_table_shared_levels counts the entries; the flags that come back are the union of what the
    levels set (exact except for qc_consistency corner cases, include/obsgrid_reports.h)."""
    case = synth.small_case(iew=60, jns=50, kbu=9, n_sfc=300, n_snd=120, radii=(5, 3), seed=104)
    assert abs(float(case.pressure[1]) - float(case.pressure[2])) * 100.0 < 2 * 1500
    rl, date, time = R.synthetic_reports(case, seed=104 * 7 + 1)
    a, b, t = _round_trip_both_ways(oracle, case, rl, date, time, 1500)
    assert R.shared_levels(t) > 0
    _assert_lists_equal(a, b)                      # this case: the union is the reference's result
    a1, b1, t1 = _round_trip_both_ways(oracle, case, rl, date, time, 1)
    assert R.shared_levels(t1) == 0
    _assert_lists_equal(a1, b1)


def test_shared_levels_are_exact_with_the_consistency_pass_left_to_the_list(oracle):
    """The recipe of include/obsgrid_reports.h for that regime: og_qc_time with run_consistency = 0, then
    og_table_to_reports (union of the levels' bits, qc_consistency once, as the reference runs it).  Seed
    179 of profiles/exp/fuzz_reports.py is the case where the per-row consistency pass changes a flag."""
    seed, p_diff = 179, 1500
    case = synth.small_case(iew=60, jns=50, kbu=int(np.random.default_rng(seed).integers(3, 12)), n_sfc=300, n_snd=120,
                            radii=(5, 3), seed=seed)
    rl, date, time = R.synthetic_reports(case, seed=seed * 7 + 1)
    a, rows, lst = rl.copy(), rl.copy(), rl.copy()
    assert oracle.qc_time_reports(case, a, date, time, p_diff) == 0
    _use_requests_of_proc_oa(oracle, case, a, date, time, p_diff)
    t = _qc_through_table(case, rows, date, time, p_diff, lambda c: oracle.qc_time(c))
    assert R.shared_levels(t) > 0
    with pytest.raises(AssertionError):
        _assert_lists_equal(a, rows)                                   # consistency per row: one flag differs
    _qc_through_table(case, lst, date, time, p_diff, lambda c: oracle.qc_time(c, run_consistency=False))
    _assert_lists_equal(a, lst)                                        # consistency on the list: exact
    # and the table rebuilt from that list is what proc_oa's 'use' requests read
    t2 = R.reports_to_table(lst.copy(), case.iew, case.jns, case.pressure, date, time, p_diff)
    for k in range(case.kbu):
        for var in (1, 2, 3, 4):
            got = oracle.get_obs(a.copy(), False, var, float(case.pressure[k]), date, time, p_diff, FAILS_ERROR_MAX,
                                 case.iew, case.jns)
            idx, v, q = _select_from_table(t2, k, var, FAILS_ERROR_MAX, False)
            np.testing.assert_array_equal(got["index"], idx, err_msg=f"k={k} var={var}")
            np.testing.assert_array_equal(got["qc"], q[idx])
