"""The compiled host side (include/obsgrid_gpu.hpp + host/obsgrid_host.cpp): it builds against the
C ABI, refuses to run without an sm_100 device (no CPU fallback), and on a B200 reproduces the
oracle's analysis time through the C++ wrapper's proc_qc / proc_oa."""
import subprocess

import numpy as np
import pytest

from obsgrid_b200 import hostio, synth


def _case():
    return synth.small_case(iew=52, jns=44, kbu=6, n_sfc=400, n_snd=90, radii=(6, 4, 2), seed=17)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_host_driver_builds_and_rejects_bad_input(tmp_path):
    exe = hostio.build_host()
    assert exe.exists()
    assert subprocess.run([str(exe)], capture_output=True).returncode == 2           # usage
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"nope")
    r = subprocess.run([str(exe), str(bad), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 2 and "not a case file" in r.stderr


def test_case_file_round_trip(tmp_path):
    """The flat binary form holds every array the C ABI needs, in the declared order."""
    case = _case()
    p = tmp_path / "case.bin"
    hostio.write_case(case, p)
    raw = p.read_bytes()
    n3, nt = case.kbu * case.iew * case.jns, case.kbu * case.nrep
    import ctypes as C
    head = 4 + 16 + 4 + C.sizeof(case.qc_params) + C.sizeof(case.oa_params)
    floats = case.kbu + 4 * n3 + case.iew * case.jns + 4 * case.nrep + 4 * nt + case.nrep
    ints = 4 * nt + case.nrep
    assert len(raw) == head + 4 * (floats + ints)
    assert np.array_equal(np.frombuffer(raw, np.float32, case.kbu, head), case.pressure)
    assert np.array_equal(np.frombuffer(raw, np.float32, n3, head + 4 * case.kbu).reshape(case.t.shape), case.t)


@pytest.mark.skipif(_has_gpu(), reason="needs a box without a GPU")
def test_host_driver_has_no_cpu_fallback(tmp_path):
    rc, res, err = hostio.run_host(_case(), tmp_path)
    assert rc == 3 and res is None
    assert "og_init" in err and ("sm_100" in err or "no CPU fallback" in err)


@pytest.mark.gpu
@pytest.mark.parametrize("options", [(), ("--one-call",)])
def test_host_driver_matches_the_oracle(tmp_path, oracle, options):
    case = _case()
    rc, res, err = hostio.run_host(case, tmp_path, *options)
    assert rc == 0, err
    cpu = case.copy()
    oracle.qc_time(cpu)
    for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp"):
        np.testing.assert_array_equal(getattr(res, nm), getattr(cpu, nm), err_msg=nm)
    oracle.oa_time(cpu)
    for nm in ("t", "slp"):
        np.testing.assert_array_equal(getattr(res, nm), getattr(cpu, nm), err_msg=nm)
    for nm in ("u", "v", "rh"):                                      # bananas: atan2 -> north-star tolerance
        np.testing.assert_allclose(getattr(res, nm), getattr(cpu, nm), rtol=1e-5, atol=1e-3, err_msg=nm)
    assert np.any(res.t != case.t) and np.count_nonzero(res.qc_t) > 0


@pytest.mark.gpu
def test_host_driver_equals_the_python_mirror(tmp_path, ctx):
    case = _case()
    rc, res, err = hostio.run_host(case, tmp_path)
    assert rc == 0, err
    py = case.copy()
    ctx.proc_qc(py); ctx.proc_oa(py)
    for nm in hostio.FIELDS + hostio.FLAGS:
        np.testing.assert_array_equal(getattr(res, nm), getattr(py, nm), err_msg=nm)
