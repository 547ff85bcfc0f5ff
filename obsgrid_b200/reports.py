"""Report lists: the plain-array image of the reference's TYPE(report) / TYPE(measurement) list
(include/obsgrid_reports.h) and its conversion to and from the flat per-time observation table.

Host-side mirror of `og_reports_to_table` / `og_table_to_reports` plus a seeded generator of report
lists that exercises every branch of query_ob (src/ob_access_module.F90:11-535): surface reports,
soundings with interpolated analysis levels and significant levels in between, single-level reports
off an analysis level, reports outside the time window, without elevation, without a surface level,
outside the domain, bogus reports, discarded ones.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import OG_MISSING_R, OgTimeDesc, check

MISSING = float(OG_MISSING_R)


class OgField(C.Structure):
    _fields_ = [("data", C.c_float), ("qc", C.c_int)]


class OgMeas(C.Structure):
    _fields_ = [(n, OgField) for n in ("pressure", "height", "temperature", "dew_point", "speed",
                                       "direction", "u", "v", "rh", "thickness")]


class OgReport(C.Structure):
    _fields_ = [("latitude", C.c_float), ("longitude", C.c_float), ("x", C.c_float), ("y", C.c_float),
                ("elevation", C.c_float), ("platform", C.c_char * 40), ("date_char", C.c_char * 14),
                ("discard", C.c_int), ("bogus", C.c_int), ("slp", OgField), ("psfc", OgField),
                ("first_meas", C.c_int), ("n_meas", C.c_int)]


MEAS_DTYPE = np.dtype([(n, [("data", np.float32), ("qc", np.int32)]) for n, _ in OgMeas._fields_])
assert MEAS_DTYPE.itemsize == C.sizeof(OgMeas)


class ReportList:
    """reports: ctypes array of OgReport; meas: numpy structured array (MEAS_DTYPE)."""

    def __init__(self, reports, meas):
        self.reports, self.meas = reports, meas

    @property
    def nrep(self):
        return len(self.reports)

    def copy(self) -> "ReportList":
        r = (OgReport * self.nrep)()
        C.memmove(r, self.reports, C.sizeof(r))
        return ReportList(r, self.meas.copy())

    def meas_ptr(self):
        return self.meas.ctypes.data_as(C.c_void_p)

    def header(self, name):
        return np.array([getattr(r, name) for r in self.reports])


class Table:
    """The observation rows of an og_time_desc, owned as numpy arrays (plus meas_index)."""

    ROWS_F = ("ob_u", "ob_v", "ob_t", "ob_rh", "ob_td")
    ROWS_I = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_td")

    def __init__(self, kbu, nrep):
        self.kbu, self.nrep = kbu, nrep
        for n in self.ROWS_F:
            setattr(self, n, np.full((kbu, nrep), MISSING, np.float32))
        for n in self.ROWS_I:
            setattr(self, n, np.zeros((kbu, nrep), np.int32))
        self.ob_slp = np.full(nrep, MISSING, np.float32)
        self.qc_slp = np.zeros(nrep, np.int32)
        self.bogus = np.zeros(nrep, np.int32)
        for n in ("xob", "yob", "lon", "elev"):
            setattr(self, n, np.zeros(nrep, np.float32))
        self.meas_index = np.full((kbu, nrep), -1, np.int32)

    NAMES = ROWS_F + ROWS_I + ("ob_slp", "qc_slp", "bogus", "xob", "yob", "lon", "elev")

    def attach(self, d: OgTimeDesc):
        d.nrep, d.kbu = self.nrep, self.kbu
        for n in self.NAMES:
            setattr(d, n, getattr(self, n).ctypes.data)
        return d


def reports_to_table(rl: ReportList, iew, jns, pressure, date, time, request_p_diff=1) -> Table:
    pressure = np.ascontiguousarray(pressure, np.float32)
    t = Table(pressure.size, rl.nrep)
    d = OgTimeDesc()
    d.iew, d.jns, d.pressure = iew, jns, pressure.ctypes.data
    t.attach(d)
    check(_capi.lib().og_reports_to_table(rl.reports, rl.nrep, rl.meas_ptr(), rl.meas.size, date, time,
                                          request_p_diff, C.byref(d), t.meas_index.ctypes.data_as(C.c_void_p)),
          "og_reports_to_table")
    t.pressure, t.iew, t.jns = pressure, iew, jns
    return t


def table_to_reports(t: Table, rl: ReportList, date, time):
    d = OgTimeDesc()
    d.iew, d.jns, d.pressure = t.iew, t.jns, t.pressure.ctypes.data
    t.attach(d)
    check(_capi.lib().og_table_to_reports(rl.reports, rl.nrep, rl.meas_ptr(), rl.meas.size, date, time, C.byref(d),
                                          t.meas_index.ctypes.data_as(C.c_void_p)), "og_table_to_reports")


def shared_levels(t: Table) -> int:
    """Table entries whose measurement also answers a lower level of the same report
    (og_table_shared_levels: 0 means the table path is exact, see include/obsgrid_reports.h)."""
    return int(_capi.lib().og_table_shared_levels(t.meas_index.ctypes.data_as(C.c_void_p), t.kbu, t.nrep))


def store_ob_defect_reports(rl: ReportList, pressure) -> list:
    """Reports that trip the reference's store_ob defect (include/obsgrid_reports.h): at least two
    measurements, and the FIRST measurement within 1 Pa of some analysis level is not the one query_ob
    answers that level with (query_ob skips a surface measurement: height within 1 m of the elevation)."""
    out = []
    p_pa = [np.float32(p) * np.float32(100.0) for p in np.asarray(pressure, np.float32) if int(round(float(p))) != 1001]
    for i, r in enumerate(rl.reports):
        if r.n_meas < 2:
            continue
        m = rl.meas[r.first_meas:r.first_meas + r.n_meas]
        for lv in p_pa:
            near = np.flatnonzero(np.abs(m["pressure"]["data"] - lv) < np.float32(1.0))
            if near.size == 0:
                continue
            first = near[0]
            h = m["height"]["data"][first]
            is_sfc = abs(float(h) - float(r.elevation)) < 1.0 and abs(float(h) - MISSING) >= 1.0
            if is_sfc and near.size > 1:
                out.append(i)
                break
    return out


def _pad(s: str, n: int) -> bytes:
    return s.encode().ljust(n)[:n]


def synthetic_reports(case, seed=0, date=20240115, time=120000):
    """A report list over the grid of `case` (a synth.TimeCase: its first guess provides plausible
    values) with every flavour query_ob distinguishes.  Returns (ReportList, date, time)."""
    rng = np.random.default_rng(seed)
    iew, jns, K = case.iew, case.jns, case.kbu
    p_pa = case.pressure.astype(np.float64) * 100.0
    n = case.nrep
    reps = (OgReport * n)()
    meas = []

    def fld(v, q=0):
        return (np.float32(v), np.int32(q))

    def level(p, h, t, td, u, v, rh):
        spd = MISSING if (u == MISSING or v == MISSING) else float(np.hypot(u, v))
        return (fld(p), fld(h), fld(t), fld(td), fld(spd), fld(MISSING), fld(u), fld(v), fld(rh), fld(MISSING))

    hh = (time // 10000)
    for r in range(n):
        R = reps[r]
        x, y = float(case.xob[r]), float(case.yob[r])
        kind = rng.choice(["sfc", "sounding", "airep", "sfc_only_slp", "no_elev"], p=[0.55, 0.25, 0.1, 0.05, 0.05])
        if rng.random() < 0.04:                      # outside the domain of the analysis
            x = float(rng.choice([1.3, iew - 0.2, x])); y = float(rng.choice([0.4, jns + 3.0])) if x == float(case.xob[r]) else y
        R.x, R.y = x, y
        R.latitude, R.longitude = 30.0 + 0.01 * y, float(case.lon[r])
        elev = float(case.elev[r])
        R.elevation = MISSING if kind == "no_elev" else elev
        plat = {"sfc": "FM-12 SYNOP", "sounding": "FM-35 TEMP", "airep": "FM-97 AIREP",
                "sfc_only_slp": "FM-13 SHIP", "no_elev": "FM-12 SYNOP"}[kind]
        p40 = list(plat.ljust(40))
        if rng.random() < 0.1:                       # time window override in platform(36:39)
            p40[35:39] = list(rng.choice(["  90", "7200", " abc", "+600"]))
        R.platform = "".join(p40).encode()
        # valid time: mostly the analysis hour, some off by minutes / hours
        dmin = int(rng.choice([0, 0, 0, 10, -25, 29, 31, -45, 59, 61, 125, -180]))
        tot = hh * 60 + dmin
        day = date % 100 + (tot // (24 * 60))
        tot %= 24 * 60
        R.date_char = _pad(f"{date // 100:06d}{day:02d}{tot // 60:02d}{tot % 60:02d}00", 14)
        R.discard = int(rng.random() < 0.02)
        R.bogus = int(rng.random() < 0.03)
        slp_here = float(case.ob_slp[r]) if rng.random() < 0.8 else MISSING
        R.slp = OgField(slp_here, int(rng.choice([0, 0, 0, 4, 1 << 16])))
        R.psfc = OgField(MISSING, 0)
        R.first_meas = len(meas)

        def vals(k):
            g = lambda a: float(a[k, r])
            t_, u_, v_, rh_ = g(case.ob_t), g(case.ob_u), g(case.ob_v), g(case.ob_rh)
            if t_ == MISSING: t_ = 280.0 - 0.05 * k
            if u_ == MISSING: u_, v_ = 5.0, -3.0
            if rh_ == MISSING: rh_ = 55.0
            drop = rng.random(4) < 0.07
            t_, u_, v_, rh_ = (MISSING if drop[0] else t_), (MISSING if drop[1] else u_), (MISSING if drop[2] else v_), (MISSING if drop[3] else rh_)
            td_ = MISSING
            if t_ != MISSING and rng.random() < 0.8:
                td_ = t_ - float(rng.uniform(-1.5, 12.0))          # some dew points above the temperature
            return t_, td_, u_, v_, rh_

        if kind in ("sfc", "no_elev"):
            t_, td_, u_, v_, rh_ = vals(0)
            h = elev + float(rng.choice([0.0, 0.0, 0.4, 0.99, 5.0]))     # height vs elevation within / beyond 1 m
            meas.append(level(float(rng.uniform(85000, 102000)), h, t_, td_, u_, v_, rh_))
        elif kind == "sounding":
            t_, td_, u_, v_, rh_ = vals(0)
            if rng.random() < 0.85:                  # a surface level first
                meas.append(level(float(rng.uniform(85000, 101000)), elev, t_, td_, u_, v_, rh_))
            ks = sorted(rng.choice(np.arange(1, K), size=min(K - 1, int(rng.integers(2, 7))), replace=False))
            for k in ks:                             # analysis levels, exact or a fraction of a Pa off
                t_, td_, u_, v_, rh_ = vals(int(k))
                off = float(rng.choice([0.0, 0.0, 0.5, -0.9, 1.5]))
                meas.append(level(p_pa[k] + off, elev + 150.0 * (k + 1), t_, td_, u_, v_, rh_))
                if rng.random() < 0.4:               # a significant level between analysis levels
                    meas.append(level(p_pa[k] - 700.0, elev + 150.0 * (k + 1) + 60.0, t_, td_, u_, v_, rh_))
        elif kind == "airep":
            k = int(rng.integers(1, K))
            t_, td_, u_, v_, rh_ = vals(k)
            off = float(rng.choice([0.0, 0.4, 300.0, -1200.0, 2600.0]))  # single level, maybe off the analysis level
            meas.append(level(p_pa[k] + off, MISSING if rng.random() < 0.5 else elev + 5000.0, t_, td_, u_, v_, rh_))
        R.n_meas = len(meas) - R.first_meas
    m = np.array(meas, dtype=MEAS_DTYPE) if meas else np.zeros(0, MEAS_DTYPE)
    # incoming flags on some measurements
    for nm in ("u", "v", "temperature", "rh"):
        sel = rng.random(m.size) < 0.05
        m[nm]["qc"][sel] = rng.choice([4, 1 << 14, 1 << 16, 1 << 17, 1 << 20], size=int(sel.sum()))
    return ReportList(reps, m), date, time
