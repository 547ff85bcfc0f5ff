"""Seeded synthetic first guess + observation tables for the five benchmark configurations
(BASELINE.json `configs`).

The generator emits the *post-reader* arrays the hot path consumes -- first-guess fields in
the reference's (jns,iew,kbu) y-fastest layout (src/first_guess.inc) and a flat per-time
observation table in report order -- so neither the benchmark nor the parity tests depend on
the Fortran met_em / little_r ingest.  PRNG: numpy PCG64 seeded with the config number.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._capi import OG_MAX_SCAN, OG_MISSING_R, OgObsCsr, OgOaParams, OgQcParams, OgTimeDesc

# name -> (iew, jns, dx_m, kbu, n_surface, n_sounding, radii, n_times)
CONFIGS = {
    1: dict(iew=100, jns=100, dx_m=30000.0, kbu=27, n_sfc=1800, n_snd=200,
            radii=(10, 8, 6, 4), times=1,
            label="cfg1: 100x100 Lambert 30 km, 27 levels, 2k reports, Cressman 10/8/6/4 + qc1/qc2"),
    2: dict(iew=500, jns=500, dx_m=12000.0, kbu=38, n_sfc=45000, n_snd=5000,
            radii=(10, 8, 6, 4), times=1,
            label="cfg2: 500x500 12 km, 38 levels, 50k reports, Cressman 10/8/6/4 + qc1/qc2"),
    3: dict(iew=1000, jns=1000, dx_m=9000.0, kbu=38, n_sfc=180000, n_snd=20000,
            radii=(10, 8, 6, 4), times=16,
            label="cfg3: 1000x1000 9 km, 38 levels, 200k reports, 16 times"),
    4: dict(iew=2000, jns=2000, dx_m=3000.0, kbu=27, n_sfc=500000, n_snd=0,
            radii=(10, 8, 6, 4), times=24, surface_only=True,
            label="cfg4: surface FDDA 2000x2000 3 km, 500k surface reports, 24 times"),
    5: dict(iew=4000, jns=4000, dx_m=3000.0, kbu=50, n_sfc=1800000, n_snd=200000,
            radii=(12, 9, 6, 4, 2), times=1,
            label="cfg5: 4000x4000 3 km, 50 levels, 2M reports, Cressman 12/9/6/4/2"),
}


def pressure_levels(kbu: int) -> np.ndarray:
    """k=0 is the surface pseudo-level 1001 hPa (first_guess_module.F90:326, driver.F90:289);
    every set contains 850/700/500 hPa (final_analysis_module.F90:1111-1127)."""
    if kbu == 27:
        lev = [1000, 975, 950, 925, 900, 850, 800, 750, 700, 650, 600, 550, 500, 450, 400, 350,
               300, 250, 200, 150, 100, 70, 50, 30, 20, 10]
    elif kbu == 38:
        lev = list(range(1000, 99, -25))
    elif kbu == 50:
        lev = sorted(list(range(1000, 99, -25)) + [987.5, 962.5, 937.5, 90, 80, 70, 60, 50, 40,
                                                    30, 20, 10], reverse=True)
    else:
        base = [1000, 850, 700, 500, 300, 200, 100]
        if kbu - 1 <= len(base):
            lev = base[:kbu - 1]
        else:
            extra = np.linspace(975, 60, kbu - 1 - len(base)).tolist()
            lev = sorted(set(base) | set(np.round(extra, 1).tolist()), reverse=True)[:kbu - 1]
    assert len(lev) == kbu - 1, (kbu, len(lev))
    return np.array([1001.0] + list(lev), np.float32)


def dx_km_like_driver(dx_m: float) -> np.float32:
    """driver.F90:222,250,265: dxd*0.001, *1000., *0.001 in REAL(4)."""
    d = np.float32(dx_m) * np.float32(0.001)
    d = np.float32(d * np.float32(1000.0))
    return np.float32(d * np.float32(0.001))


def lambert_lon(x, y, iew, jns, dx_m, truelat1=30.0, truelat2=60.0, stdlon=-98.0, lat0=40.0):
    """Longitude (deg) of dot-grid point (x,y) on a Lambert conformal grid centred at
    (lat0, stdlon): inverse of llij_lc (map_utils_module.F90:703-758).  Only feeds local_time."""
    rad = np.pi / 180.0
    re = 6370000.0
    cone = (np.log(np.cos(truelat1 * rad)) - np.log(np.cos(truelat2 * rad))) / (
        np.log(np.tan((90.0 - abs(truelat1)) * rad * 0.5)) -
        np.log(np.tan((90.0 - abs(truelat2)) * rad * 0.5)))
    rebydx = re / dx_m
    ctl1r = np.cos(truelat1 * rad)
    rsw = rebydx * ctl1r / cone * (np.tan((90.0 - lat0) * rad / 2.0) /
                                   np.tan((90.0 - truelat1) * rad / 2.0)) ** cone
    polei = (iew + 1) / 2.0
    polej = (jns + 1) / 2.0 + rsw
    xx = np.asarray(x, np.float64) - polei
    yy = polej - np.asarray(y, np.float64)
    lon = stdlon + np.degrees(np.arctan2(xx, yy)) / cone
    lon = np.where(lon > 180.0, lon - 360.0, lon)
    lon = np.where(lon < -180.0, lon + 360.0, lon)
    return lon.astype(np.float32)


def _bilinear64(g_xy, x, y):
    """g_xy: (jns, iew) array [j, i]; x,y 1-based coords."""
    i = np.floor(x).astype(np.int64); j = np.floor(y).astype(np.int64)
    dx = x - i; dy = y - j
    i0 = i - 1; j0 = j - 1
    return ((1 - dx) * ((1 - dy) * g_xy[j0, i0] + dy * g_xy[j0 + 1, i0]) +
            dx * ((1 - dy) * g_xy[j0, i0 + 1] + dy * g_xy[j0 + 1, i0 + 1]))


@dataclass
class TimeCase:
    """One analysis time: first guess, observation table, namelist parameters."""
    iew: int
    jns: int
    kbu: int
    dx_km: np.float32
    pressure: np.ndarray
    t: np.ndarray      # (kbu, iew, jns) C order == Fortran (jns,iew,kbu)
    u: np.ndarray
    v: np.ndarray
    rh: np.ndarray
    slp: np.ndarray    # (iew, jns) hPa
    xob: np.ndarray
    yob: np.ndarray
    lon: np.ndarray
    elev: np.ndarray
    ob_u: np.ndarray   # (kbu, nrep)
    ob_v: np.ndarray
    ob_t: np.ndarray
    ob_rh: np.ndarray
    ob_slp: np.ndarray  # (nrep,) Pa
    qc_u: np.ndarray
    qc_v: np.ndarray
    qc_t: np.ndarray
    qc_rh: np.ndarray
    qc_slp: np.ndarray
    qc_params: OgQcParams = field(default_factory=OgQcParams)
    oa_params: OgOaParams = field(default_factory=OgOaParams)
    label: str = ""
    # optional rows of og_time_desc (ABI 3): dew point of the measurements and its flag (kbu, nrep),
    # bogus reports (nrep,)
    ob_td: np.ndarray | None = None
    qc_td: np.ndarray | None = None
    bogus: np.ndarray | None = None

    @property
    def nrep(self) -> int:
        return int(self.xob.size)

    _FIELDS = ("t", "u", "v", "rh", "slp")
    _FLAGS = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")

    def desc(self) -> OgTimeDesc:
        d = OgTimeDesc()
        d.iew, d.jns, d.kbu, d.dx_km = self.iew, self.jns, self.kbu, float(self.dx_km)
        d.nrep = self.nrep
        for name in ("pressure", "t", "u", "v", "rh", "slp", "xob", "yob", "lon", "elev", "ob_u",
                     "ob_v", "ob_t", "ob_rh", "ob_slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
            a = getattr(self, name)
            assert a.flags["C_CONTIGUOUS"], name
            setattr(d, name, a.ctypes.data)
        for name in ("ob_td", "qc_td", "bogus"):
            a = getattr(self, name)
            if a is not None:
                assert a.flags["C_CONTIGUOUS"], name
                setattr(d, name, a.ctypes.data)
        return d

    def copy(self) -> "TimeCase":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.copy() if isinstance(v, np.ndarray) else v
        c = TimeCase(**kw)
        qp = OgQcParams(); C.memmove(C.byref(qp), C.byref(self.qc_params), C.sizeof(qp))
        op = OgOaParams(); C.memmove(C.byref(op), C.byref(self.oa_params), C.sizeof(op))
        c.qc_params, c.oa_params = qp, op
        return c

    def slab_xy(self, name: str, k: int = 0) -> np.ndarray:
        """(jns, iew) numpy view-copy of one level == Fortran gridded(iew,jns) memory."""
        f = getattr(self, name)
        a = f if f.ndim == 2 else f[k]
        return np.ascontiguousarray(a.T)

    def field_bytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in self._FIELDS)

    def table_bytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in
                   ("xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp") +
                   self._FLAGS)


class CsrObs:
    """The observation rows of a TimeCase in sparse form (og_obs_csr): per level the reports that
    have any value there, ascending.  `alloc` maps a numpy array to the array to keep (e.g. a pinned copy)."""

    ROWS = ("u", "v", "t", "rh", "td")

    def __init__(self, case: "TimeCase", alloc=lambda a: a):
        K, n = case.kbu, case.nrep
        miss = np.float32(OG_MISSING_R)
        present = np.zeros((K, n), bool)
        for nm in ("ob_u", "ob_v", "ob_t", "ob_rh"):
            present |= np.abs(getattr(case, nm) - miss) > 1.0
        has_td = case.ob_td is not None and case.qc_td is not None
        if has_td:
            present |= np.abs(case.ob_td - miss) > 1.0
        for nm in ("qc_u", "qc_v", "qc_t", "qc_rh"):
            present |= getattr(case, nm) != 0
        k_idx, r_idx = np.nonzero(present)                   # row-major: level, then ascending report
        self.kbu, self.nrep, self.has_td = K, n, has_td
        self.lev_start = alloc(np.concatenate([[0], np.cumsum(present.sum(axis=1))]).astype(np.int32))
        self.rep = alloc(r_idx.astype(np.int32))
        self._k = k_idx
        self.nnz = int(r_idx.size)
        for nm in self.ROWS:
            if nm == "td" and not has_td:
                setattr(self, "td", None); setattr(self, "qc_td", None)
                continue
            setattr(self, nm, alloc(np.ascontiguousarray(getattr(case, "ob_" + nm)[k_idx, r_idx])))
            setattr(self, "qc_" + nm, alloc(np.ascontiguousarray(getattr(case, "qc_" + nm)[k_idx, r_idx])))

    def struct(self) -> OgObsCsr:
        s = OgObsCsr()
        s.nnz = self.nnz
        s.lev_start, s.rep = self.lev_start.ctypes.data, self.rep.ctypes.data
        for nm in self.ROWS:
            a, q = getattr(self, nm), getattr(self, "qc_" + nm)
            setattr(s, nm, a.ctypes.data if a is not None else None)
            setattr(s, "qc_" + nm, q.ctypes.data if q is not None else None)
        return s

    def scatter_flags(self, case: "TimeCase"):
        """Write the (downloaded) compact flag rows into the dense rows of `case`."""
        for nm in self.ROWS:
            q = getattr(self, "qc_" + nm)
            if q is not None:
                getattr(case, "qc_" + nm)[self._k, self.rep] = q
        if self.has_td:
            case.ob_td[self._k, self.rep] = self.td

    def nbytes_up(self) -> int:
        return sum(a.nbytes for a in [self.lev_start, self.rep] + [getattr(self, n) for n in self.ROWS if getattr(self, n) is not None]
                   + [getattr(self, "qc_" + n) for n in self.ROWS if getattr(self, "qc_" + n) is not None])


def default_qc_params(time_hhmmss: int = 120000) -> OgQcParams:
    """namelist.oa &record4 of the reference (namelist.oa:28-50)."""
    p = OgQcParams()
    p.qc_test_error_max = 1; p.qc_test_buddy = 1
    p.max_error_t = 10; p.max_error_uv = 13; p.max_error_rh = 50
    p.max_error_dewpoint = 20; p.max_error_p = 600
    p.max_buddy_t = 8; p.max_buddy_uv = 8; p.max_buddy_rh = 40
    p.max_buddy_dewpoint = 20; p.max_buddy_p = 800
    p.buddy_weight = 1.0           # driver.F90:456 literal
    p.time_hhmmss = time_hhmmss
    p.qc_dewpoint = 1
    return p


def default_oa_params(radii=(10, 8, 6, 4), surface_only=False, scale_rh=False,
                      sfc_mult=1.0) -> OgOaParams:
    p = OgOaParams()
    for i in range(OG_MAX_SCAN):
        p.radius_influence[i] = radii[i] if i < len(radii) else -1
    p.radius_influence_sfc_mult = sfc_mult
    p.use_first_guess = 1
    p.scale_cressman_rh_decreases = int(scale_rh)
    p.smooth_type = 1
    p.surface_only = int(surface_only)
    p.clamp_fg_rh = 1
    return p


def first_guess(iew: int, jns: int, pressure: np.ndarray, phase: float = 0.0, levels=None):
    """Analytic atmosphere: lapse-rate T with waves, a meandering zonal jet peaking 60 m/s at
    250 hPa over a weak background (so circle / ellipse / banana obs all occur aloft), RH
    30-95 % with gradients, SLP 1000-1025 hPa waves.  Returns float32 (kbu, iew, jns); with
    `levels` only those levels of the configuration, in that order (same values)."""
    kbu = pressure.size
    levels = list(range(kbu)) if levels is None else list(levels)
    i = np.arange(1, iew + 1, dtype=np.float64)[:, None]
    j = np.arange(1, jns + 1, dtype=np.float64)[None, :]
    X = i / iew; Y = j / jns
    t = np.empty((len(levels), iew, jns), np.float32); u = np.empty_like(t); v = np.empty_like(t)
    rh = np.empty_like(t)
    amp = 0.09 * jns; kw = 2 * np.pi * 3.0 / iew
    yc = jns / 2.0 + amp * np.sin(kw * i + phase)
    dyc = amp * kw * np.cos(kw * i + phase)
    width = 0.11 * jns
    for slot, k in enumerate(levels):
        p = min(float(pressure[k]), 1000.0)
        tk = 288.0 * (p / 1000.0) ** 0.19 + 8.0 * np.sin(2 * np.pi * X + phase) * np.cos(
            2 * np.pi * Y) - 15.0 * (Y - 0.5)
        if k == 0:
            tk = tk + 1.0
        jet = 8.0 + 52.0 * np.exp(-((np.log(p / 250.0)) / 0.75) ** 2)
        core = np.exp(-((j - yc) / width) ** 2)
        uk = 3.0 + jet * core
        vk = (jet * core) * dyc + 1.5 * np.sin(2 * np.pi * (X + Y))
        if k == 0:
            uk = uk * 0.6; vk = vk * 0.6
        rhk = 62.0 + 30.0 * np.sin(2 * np.pi * (X * 1.5 + 0.3 * k / kbu) + phase) * np.cos(
            2 * np.pi * Y * 1.2) - 20.0 * (1.0 - p / 1000.0)
        t[slot] = tk; u[slot] = uk; v[slot] = vk; rh[slot] = np.clip(rhk, 5.0, 98.0)
    slp = (1012.5 + 10.0 * np.sin(2 * np.pi * X * 1.3 + phase) * np.cos(2 * np.pi * Y) +
           2.5 * np.cos(2 * np.pi * Y * 2.0)).astype(np.float32)
    return t, u, v, rh, slp


def make_case(cfg: int | dict, seed: int | None = None, time_index: int = 0,
              n_sfc: int | None = None, n_snd: int | None = None, kbu: int | None = None,
              gross_frac: float = 0.02, scale_rh: bool = False, sfc_mult: float = 1.0,
              sort_reports: bool = False, levels=None) -> TimeCase:
    """Build one analysis time of configuration `cfg` (1..5, or a dict like CONFIGS[n]).

    levels: an increasing list of level indices starting with 0 -> a TimeCase that holds only those
    levels of the configuration, with exactly the values the full case has there (the random stream
    is advanced through the skipped levels): the parity tests of the largest configurations check
    sampled levels without building 50 levels of 4000 x 4000 points."""
    c = dict(CONFIGS[cfg]) if not isinstance(cfg, dict) else dict(cfg)
    if n_sfc is not None: c["n_sfc"] = n_sfc
    if n_snd is not None: c["n_snd"] = n_snd
    if kbu is not None: c["kbu"] = kbu
    seed = (cfg if isinstance(cfg, int) else 0) if seed is None else seed
    rng = np.random.Generator(np.random.PCG64([seed, time_index]))
    iew, jns, K = c["iew"], c["jns"], c["kbu"]
    pressure = pressure_levels(K)
    phase = 0.4 * time_index
    if levels is not None:
        levels = [int(k) for k in levels]
        assert levels and levels[0] == 0 and all(a < b for a, b in zip(levels, levels[1:])) and levels[-1] < K
    lev = list(range(K)) if levels is None else levels
    slot_of = {k: s_ for s_, k in enumerate(lev)}
    t, u, v, rh, slp = first_guess(iew, jns, pressure, phase, lev)

    nrep = c["n_sfc"] + c["n_snd"]
    # station positions: 80 % uniform, 20 % in 2-D Gaussian clusters; strictly inside
    # 2 <= x <= iew-1 (proc_ob_access.F90:231-234)
    n_cl = int(0.2 * nrep)
    xs = rng.uniform(2.0, iew - 1.0, nrep); ys = rng.uniform(2.0, jns - 1.0, nrep)
    if n_cl > 0:
        ncent = 8
        cx = rng.uniform(0.15 * iew, 0.85 * iew, ncent); cy = rng.uniform(0.15 * jns, 0.85 * jns, ncent)
        which = rng.integers(0, ncent, n_cl)
        xs[:n_cl] = cx[which] + rng.normal(0, 0.03 * iew, n_cl)
        ys[:n_cl] = cy[which] + rng.normal(0, 0.03 * jns, n_cl)
        xs = np.clip(xs, 2.0, iew - 1.0); ys = np.clip(ys, 2.0, jns - 1.0)
    perm = rng.permutation(nrep)
    xs, ys = xs[perm], ys[perm]
    if sort_reports:    # sort_obs orders reports by location (obs_sort_module.F90 merge sort)
        o = np.lexsort((ys, xs)); xs, ys = xs[o], ys[o]
    xob = xs.astype(np.float32); yob = ys.astype(np.float32)
    is_snd = np.zeros(nrep, bool)
    if c["n_snd"] > 0:
        is_snd[rng.choice(nrep, c["n_snd"], replace=False)] = True
    lon = lambert_lon(xob, yob, iew, jns, c["dx_m"])
    elev = rng.uniform(0.0, 2500.0, nrep).astype(np.float32)

    x64 = xob.astype(np.float64); y64 = yob.astype(np.float64)
    sig = dict(u=2.0, v=2.0, t=1.0, rh=8.0, slp=150.0)

    def observe(g_ij, sigma, present, lo=None, hi=None, scale=1.0):
        # truth = first guess + smooth large-scale perturbation; ob = truth + noise (+ gross)
        if g_ij is None:      # a level that is not kept: only advance the random stream
            rng.normal(0.0, sigma, nrep); rng.random(nrep)
            rng.choice([-1.0, 1.0], nrep); rng.uniform(5, 20, nrep)
            return None
        truth = _bilinear64(g_ij.T.astype(np.float64), x64, y64) * scale
        truth = truth + 1.5 * sigma * np.sin(2 * np.pi * x64 / iew * 2.0 + phase) * np.cos(
            2 * np.pi * y64 / jns * 2.0)
        ob = truth + rng.normal(0.0, sigma, nrep)
        gross = rng.random(nrep) < gross_frac
        ob = np.where(gross, ob + rng.choice([-1.0, 1.0], nrep) * rng.uniform(5, 20, nrep) * sigma, ob)
        if lo is not None:
            ob = np.clip(ob, lo, hi)
        out = np.full(nrep, OG_MISSING_R, np.float32)
        out[present] = ob[present].astype(np.float32)
        return out

    KS = len(lev)
    ob_u = np.empty((KS, nrep), np.float32); ob_v = np.empty_like(ob_u)
    ob_t = np.empty_like(ob_u); ob_rh = np.empty_like(ob_u)
    everyone = np.ones(nrep, bool)
    for k in range(K):
        present = everyone if k == 0 else is_snd
        s_ = slot_of.get(k)
        if s_ is None:
            for nm in ("u", "v", "t", "rh"):
                observe(None, sig[nm], present)
            continue
        ob_u[s_] = observe(u[s_], sig["u"], present)
        ob_v[s_] = observe(v[s_], sig["v"], present)
        ob_t[s_] = observe(t[s_], sig["t"], present)
        ob_rh[s_] = observe(rh[s_], sig["rh"], present, 1.0, 100.0)
    ob_slp = observe(slp, sig["slp"], everyone, scale=100.0)
    pressure = pressure[lev]
    K = KS

    z = lambda shape: np.zeros(shape, np.int32)
    case = TimeCase(iew=iew, jns=jns, kbu=K, dx_km=dx_km_like_driver(c["dx_m"]), pressure=pressure,
                    t=t, u=u, v=v, rh=rh, slp=slp, xob=xob, yob=yob, lon=lon, elev=elev,
                    ob_u=ob_u, ob_v=ob_v, ob_t=ob_t, ob_rh=ob_rh, ob_slp=ob_slp,
                    qc_u=z((K, nrep)), qc_v=z((K, nrep)), qc_t=z((K, nrep)), qc_rh=z((K, nrep)),
                    qc_slp=z(nrep),
                    qc_params=default_qc_params(int(((12 + time_index * 6) % 24) * 10000)),
                    oa_params=default_oa_params(c["radii"], c.get("surface_only", False),
                                                scale_rh, sfc_mult),
                    label=c.get("label", "custom"))
    return case


def small_case(iew=40, jns=36, kbu=5, n_sfc=150, n_snd=40, seed=7, dx_m=30000.0,
               radii=(6, 4, 2), **kw) -> TimeCase:
    """A small custom configuration for fast parity tests."""
    cfg = dict(iew=iew, jns=jns, dx_m=dx_m, kbu=kbu, n_sfc=n_sfc, n_snd=n_snd, radii=radii,
               times=1, label=f"small {iew}x{jns}x{kbu}")
    return make_case(cfg, seed=seed, **kw)
