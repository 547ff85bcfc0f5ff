"""obsgrid_b200/perslab.py -- the Python mirror of the reference's level x variable loops
(src/proc_qc.F90:243-408, src/proc_oa.F90:356-849) over per-slab calls -- driven with the CPU oracle's
per-slab functions, against the oracle's own C restatement of the same loops (ogo_qc_time /
ogo_oa_time).  Two independent writings of the loops, the same arithmetic underneath: everything must
agree bit for bit.  (On the GPU the same mirror runs over the library's per-slab entry points:
tests/test_gpu_parity.py::test_per_slab_drop_ins_equal_the_batched_time.)"""
import numpy as np
import pytest

from obsgrid_b200 import perslab, synth


class _OracleCtx:
    def __init__(self, o):
        self.o = o
        for nm in ("local_time", "error_max", "buddy_check", "innovation", "cressman"):
            setattr(self, nm, getattr(o, nm))

    def selftest_logf(self, x):
        return self.o.logf(x)

    def qc_consistency(self, case, k_begin=0, k_end=None):
        self.o.qc_consistency(case, k_begin, k_end)


@pytest.mark.parametrize("scale_rh,sfc_mult,seed,tests", [(False, 1.0, 1, (1, 1)), (True, 0.6, 2, (1, 1)),
                                                           (False, 1.0, 3, (1, 0)), (True, 1.0, 4, (0, 1)),
                                                           (False, 0.3, 5, (0, 0))])
def test_python_loops_equal_the_c_loops(oracle, scale_rh, sfc_mult, seed, tests):
    base = synth.small_case(iew=44, jns=38, kbu=5, n_sfc=260, n_snd=70, radii=(6, 4, 2), seed=seed,
                            scale_rh=scale_rh, sfc_mult=sfc_mult)
    base.qc_params.qc_test_error_max, base.qc_params.qc_test_buddy = tests
    rng = np.random.default_rng(seed)
    base.qc_t[1, ::9] |= 1 << 16                 # incoming flags: some obs already failed, some beyond 2**20
    base.qc_u[0, ::11] |= 1 << 20
    for nm in ("ob_td", "qc_td", "bogus"):
        if getattr(base, nm, None) is None:
            break
    else:
        base.bogus[rng.random(base.nrep) < 0.05] = 1
    a, b = base.copy(), base.copy()
    oracle.qc_time(a)
    calls = perslab.proc_qc_per_slab(_OracleCtx(oracle), b)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)
    if base.ob_td is not None and base.qc_td is not None:
        np.testing.assert_array_equal(a.qc_td, b.qc_td)
        np.testing.assert_array_equal(a.ob_td, b.ob_td)
    rc = None
    try:
        oracle.oa_time(a)
    except Exception as e:          # noqa: BLE001  (the STOP of proc_oa.F90:694-703 for a tiny surface radius)
        rc = e
    if rc is None:
        calls += perslab.proc_oa_per_slab(_OracleCtx(oracle), b)
        for nm in ("t", "u", "v", "rh", "slp"):
            np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)
        assert np.any(a.t != base.t)
    assert calls >= 1
