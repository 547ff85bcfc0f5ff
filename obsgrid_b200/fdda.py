"""The surface-FDDA chain of the reference's driver (src/driver.F90:107-113, :306-341, :455-500,
:561-563) composed from the library's entry points, fields resident in HBM between the phases.

With f4d = .TRUE. the driver, for every traditional analysis time icount (interval apart):

  1. analyses the 3-D time (proc_qc + proc_oa on every level) and keeps the analysed t, u, v, rh
     and sea-level pressure (store_fa, tinterp_module.F90:10-49; PMSL in Pa);
  2. from the second time on, for every FDDA time in between (intf4d apart, interval / intf4d - 1 of
     them): temporal_interp of the two analysed times around it (tinterp_module.F90:53-215, integer
     weights) gives the first guess; proc_qc on the FDDA time's observations with the observation
     density / distance fields; proc_oa of the surface level only (fdda_loop > 1, proc_oa.F90:365).

The observation density tobbox and the distance odis start from zero at every analysis time
(driver.F90:140-141, inside fdda_ish_loop); proc_qc accumulates the five surface variables into
tobbox, rescales it and derives odis (proc_qc.F90:349-351, :412-420) -- they are outputs for the
writer (driver.F90:645, :661-666), nothing downstream of the path reads them.
Only the surface level of an FDDA time is analysed, so the chain interpolates and checks level 0 of
the 3-D fields (levels are independent slabs).

`run_chain` is written against a small backend interface: `DeviceBackend` (GPU, device pointers,
nothing leaves HBM) is what bench.py times for BASELINE configs[3]; the parity test runs the same
chain with the CPU oracle as the backend (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import numpy as np

FIELDS3 = ("t", "u", "v", "rh")


def fdda_times(n_3d: int, ratio: int):
    """(icount, icount_fdda, icount_1, icount_2) of every FDDA surface analysis, in the driver's order."""
    out = []
    for icount in range(2, n_3d + 1):
        loop_max = 1 + ratio
        c1, c2 = (icount - 2) * ratio + 1, (icount - 1) * ratio + 1
        for loop in range(2, loop_max):                     # the last pass of fdda_ish_loop is skipped (GOTO 101)
            out.append((icount, (icount - 1) * ratio + 1 - (loop_max - loop), c1, c2))
    return out


def run_chain(backend, cases3d, hourly_cases, ratio=6):
    """cases3d: the TimeCases of the traditional analysis times, in order (analysed in place by the
    backend); hourly_cases: for every entry of fdda_times(), a one-level (kbu == 1) TimeCase holding
    that FDDA time's surface observations -- its fields are overwritten with the interpolated first
    guess and then analysed.  Returns the list of analysed hourly cases (the same objects)."""
    times = fdda_times(len(cases3d), ratio)
    assert len(hourly_cases) == len(times)
    tob = backend.zeros2d(cases3d[0])
    stored = {}
    h = 0
    for icount, case in enumerate(cases3d, start=1):
        backend.clear2d(tob)                                     # driver.F90:140-141
        tob = backend.analyse_3d(case, tob)                      # fdda_loop == 1
        stored[icount] = backend.store_fa(case)                  # level 0 of t, u, v, rh; PMSL * 100
        stored.pop(icount - 2, None)
        while h < len(times) and times[h][0] == icount:
            _, cf, c1, c2 = times[h]
            hc = hourly_cases[h]
            backend.first_guess(hc, stored[icount - 1], stored[icount], cf, c1, c2)
            backend.clear2d(tob)
            tob = backend.analyse_surface(hc, tob)
            h += 1
    return hourly_cases


class HostBackend:
    """Host arrays through a `proc` object with the per-time calls (api.Context, or the oracle
    wrapped by tests): proc_qc_fdda(case, tobbox) -> (tobbox, odis), proc_oa(case),
    temporal_interp(a, b, cf, c1, c2, cross) -> array."""

    def __init__(self, proc):
        self.p = proc

    def zeros2d(self, case):
        return np.zeros((case.iew, case.jns), np.float32)

    def clear2d(self, tob):
        tob[...] = 0

    def analyse_3d(self, case, tob):
        tob, _ = self.p.proc_qc_fdda(case, tob)
        self.p.proc_oa(case)
        return tob

    def store_fa(self, case):
        s = {nm: getattr(case, nm)[0].copy() for nm in FIELDS3}
        s["slp"] = (case.slp * np.float32(100.0)).astype(np.float32)      # tinterp_module.F90:45
        return s

    def first_guess(self, hc, a, b, cf, c1, c2):
        for nm in FIELDS3:
            getattr(hc, nm)[0] = self.p.temporal_interp(a[nm], b[nm], cf, c1, c2, nm in ("t", "rh"))
        slp = self.p.temporal_interp(a["slp"], b["slp"], cf, c1, c2, True)
        slp[:-1, :-1] = slp[:-1, :-1] * np.float32(0.01)                   # driver.F90:322 / :336
        hc.slp[...] = slp

    def analyse_surface(self, hc, tob):
        tob, _ = self.p.proc_qc_fdda(hc, tob)
        self.p.proc_oa(hc)
        return tob


class DeviceBackend:
    """Everything stays on the GPU: the cases' arrays are torch tensors on the context's device (a dict
    name -> tensor per case, `dev`), descriptors carry device pointers."""

    def __init__(self, ctx, torch, device):
        self.ctx, self.torch, self.device = ctx, torch, device
        self._dev = {}
        # torch's own kernels (clone, zeros) must queue on the library's stream: the chain is one
        # stream-ordered sequence, nothing synchronises in between
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def upload(self, case):
        names = ("t", "u", "v", "rh", "slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "xob", "yob", "lon", "elev",
                 "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")
        t = {nm: self.torch.from_numpy(getattr(case, nm)).to(self.device) for nm in names}
        d = case.desc()
        for nm in names:
            setattr(d, nm, t[nm].data_ptr())
        self._dev[id(case)] = (t, d)
        return t

    def tensors(self, case):
        return self._dev[id(case)][0]

    def zeros2d(self, case):
        with self.torch.cuda.stream(self.stream):
            self.odis = self.torch.zeros((case.iew, case.jns), dtype=self.torch.float32, device=self.device)
            return self.torch.zeros((case.iew, case.jns), dtype=self.torch.float32, device=self.device)

    def clear2d(self, tob):
        with self.torch.cuda.stream(self.stream):
            tob.zero_()
            self.odis.zero_()

    def analyse_3d(self, case, tob):
        t, d = self._dev[id(case)]
        self.ctx.dev_proc_qc_fdda(d, case.qc_params, tob.data_ptr(), self.odis.data_ptr())
        self.ctx.dev_proc_oa(d, case.oa_params, 0, case.kbu)
        return tob

    def store_fa(self, case):
        t, _ = self._dev[id(case)]
        with self.torch.cuda.stream(self.stream):
            s = {nm: t[nm][0].clone() for nm in FIELDS3}
            s["slp"] = t["slp"].clone()
        self.ctx.dev_scale_field(s["slp"].data_ptr(), case.jns, case.iew, 100.0, False)
        return s

    def first_guess(self, hc, a, b, cf, c1, c2):
        t, _ = self._dev[id(hc)]
        for nm in FIELDS3:
            self.ctx.dev_temporal_interp(a[nm].data_ptr(), b[nm].data_ptr(), t[nm].data_ptr(), hc.jns, hc.iew, 1,
                                         cf, c1, c2, nm in ("t", "rh"))
        self.ctx.dev_temporal_interp(a["slp"].data_ptr(), b["slp"].data_ptr(), t["slp"].data_ptr(), hc.jns, hc.iew, 1,
                                     cf, c1, c2, True)
        self.ctx.dev_scale_field(t["slp"].data_ptr(), hc.jns, hc.iew, 0.01, True)

    def analyse_surface(self, hc, tob):
        t, d = self._dev[id(hc)]
        self.ctx.dev_proc_qc_fdda(d, hc.qc_params, tob.data_ptr(), self.odis.data_ptr())
        self.ctx.dev_proc_oa(d, hc.oa_params, 0, 1)
        return tob

    def download(self, case):
        t, _ = self._dev[id(case)]
        for nm in ("t", "u", "v", "rh", "slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
            getattr(case, nm)[...] = t[nm].cpu().numpy()


def synthetic_chain(cfg, n_3d, ratio=6, **small):
    """3-D cases and one-level hourly cases of configuration `cfg` (synth.make_case), or of a small
    custom grid when keyword arguments for synth.small_case are given."""
    from . import synth
    mk3 = (lambda ti: synth.small_case(seed=100 + ti, **small)) if small else (lambda ti: synth.make_case(cfg, time_index=ti))
    cases3d = [mk3(ti) for ti in range(n_3d)]
    for c in cases3d:
        c.oa_params.surface_only = 0          # the traditional analysis of a 3-D time covers every level
    hourly = []
    for h, _ in enumerate(fdda_times(n_3d, ratio)):
        if small:
            full = synth.small_case(seed=500 + h, **small)
            hc = _level0(full)
        else:
            hc = synth.make_case(cfg, time_index=1000 + h, levels=[0])
        hc.oa_params.surface_only = 1
        hourly.append(hc)
    return cases3d, hourly


def _level0(case):
    """A kbu == 1 TimeCase holding level 0 of `case`."""
    from . import synth
    kw = {}
    for k, v in case.__dict__.items():
        if isinstance(v, np.ndarray) and v.ndim >= 2 and v.shape[0] == case.kbu and k not in ("slp",):
            kw[k] = np.ascontiguousarray(v[:1])
        elif isinstance(v, np.ndarray):
            kw[k] = v.copy()
        else:
            kw[k] = v
    kw["kbu"] = 1
    kw["pressure"] = case.pressure[:1].copy()
    c = synth.TimeCase(**kw)
    return c.copy()
