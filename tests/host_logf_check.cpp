// Host build of the product's logarithm (obsgrid_b200/csrc/og_logf.h) for the CPU test suite:
// tests/test_logf.py compiles this file with g++ and compares it with the running libm's logf.
#include "../obsgrid_b200/csrc/og_logf.h"

extern "C" void check_logf_eval(const float* x, long long n, float* out)
{
    for (long long i = 0; i < n; ++i) out[i] = og::logf_glibc(x[i]);
}

// Every float with bit pattern in [lo, hi): number of inputs where the restatement and libm's
// logf differ (compared bit for bit); *first receives the first offending bit pattern.
extern "C" long long check_logf_sweep(unsigned lo, unsigned hi, unsigned stride, unsigned* first)
{
    long long bad = 0;
    for (unsigned long long b = lo; b < hi; b += stride) {
        const unsigned u = (unsigned)b;
        float x, a, c;
        memcpy(&x, &u, 4);
        a = logf(x);
        c = og::logf_glibc(x);
        if (memcmp(&a, &c, 4) != 0) { if (!bad && first) *first = u; ++bad; }
    }
    return bad;
}
