// og_logf.h -- single-precision natural logarithm that reproduces glibc's logf bit for bit.
//
// Why: the reference's DEWPOINT quantities go through the Fortran LOG intrinsic
// (proc_qc.F90:315-319, ob_access_module.F90:476-486), which gfortran lowers to a call of glibc's
// logf.  QC flags must be bit-identical, and a flag flips when |ob - first guess| crosses its
// threshold by one ulp -- so "a" correctly rounded logarithm is not good enough: glibc's logf is
// NOT correctly rounded (0.818 ulp), it is a specific algorithm.  That algorithm is not part of
// /root/reference; it is the third-party dependency of this one operation:
//
//   glibc >= 2.28, sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c (this image: glibc 2.39), i.e.
//   ARM Optimized Routines' logf (Szabolcs Nagy, 2017): with OFF = 0x3f330000 split
//   x = 2^k * z, z in [OFF, 2*OFF); look up c ~ the centre of z's 1/16 sub-interval,
//   r = z/c - 1 through a table of {1/c, log(c)}; log(x) = log1p(r) + log(c) + k*ln2 with a
//   degree-3 polynomial for log1p(r) - r, everything in double precision, rounded once at the end.
//
// The published table and coefficients are restated below (checked against the running libm's
// .rodata by tests/test_logf.py, and the function against libm's logf on every float in the
// argument range the path produces).  On x86-64 with FMA glibc dispatches (ifunc) to a build of
// that source in which the compiler has contracted the five multiply-adds; the expression below
// is that contracted form (disassembly of __logf_fma in glibc 2.39).  The non-contracted build
// differs only where the double-precision result lies within ~1e-16 of a rounding boundary of the
// float result (about one input in 10^9).
//
// Host and device: plain C++, fma() is DFMA on the device and a correctly rounded fma on the host.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OG_HD __host__ __device__ __forceinline__
#else
#define OG_HD static inline
#endif

namespace og {

// {invc, logc} x 16, then ln2, then A[0..2]  (glibc __logf_data, LOGF_TABLE_BITS = 4, LOGF_POLY_ORDER = 4)
#define OG_LOGF_DATA_INIT                                                                   \
    {                                                                                       \
        0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, \
        0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,  \
        0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3,    \
        0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, \
        0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0, 0x0p+0,                              \
        0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4,    \
        0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3,    \
        0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2, 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2,   \
        0x1.62e42fefa39efp-1, -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2  \
    }

#if defined(__CUDACC__)
static __constant__ double k_logf_data_dev[36] = OG_LOGF_DATA_INIT;
#endif
static const double k_logf_data_host[36] = OG_LOGF_DATA_INIT;

OG_HD const double* logf_data()
{
#if defined(__CUDA_ARCH__)
    return k_logf_data_dev;
#else
    return k_logf_data_host;
#endif
}

// Valid for finite normal x > 0 (the path only takes logs of rh/100 >= 1e-6, of pressure and of
// temperature ratios); zero, negative, subnormal, inf and nan arguments are forwarded to the
// platform's logf, which is what glibc's own special-case branch computes for them.
OG_HD float logf_glibc(float x)
{
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return logf(x);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
    memcpy(&zf, &iz, 4);
    const double* T = logf_data();
    const double invc = T[2 * i], logc = T[2 * i + 1];
    const double ln2 = T[32], a0 = T[33], a1 = T[34], a2 = T[35];
    const double z = (double)zf;
    const double r = fma(z, invc, -1.0);
    const double y0 = fma((double)k, ln2, logc);
    const double r2 = r * r;
    double y = fma(a1, r, a2);
    y = fma(a0, r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}

} // namespace og
