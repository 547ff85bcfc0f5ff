"""obsgrid_b200 -- B200-native (sm_100a) objective-analysis hot path of wrf-model/OBSGRID.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI declared in
include/obsgrid_gpu.h), `_capi` (ctypes declarations), `api` (host-side mirror of the
reference procedures cressman / error_max / buddy_check / proc_qc / proc_oa) and `synth`
(seeded synthetic first guess + observation tables).  There is no CPU fallback.
"""
__version__ = "0.1.0"
