// og_capi_util.h -- helpers shared by the two halves of the extern "C" surface (og_capi.cu,
// og_time.cu).  Internal.
#pragma once
#include "og_ctx.h"

#include <math.h>
#include <algorithm>

namespace og {

// Every entry point makes the context's device current first: the caller may be a host thread
// that has never touched CUDA (its current device is 0) while the context lives on another GPU.
#define OG_BIND(c)                                                                          \
    do {                                                                                      \
        if ((c) && cudaSetDevice((c)->device) != cudaSuccess) {                               \
            og::set_error("%s: cudaSetDevice(%d) failed", __func__, (c)->device);                \
            return OG_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define REQUIRE(cond, msg)                                                                \
    do { if (!(cond)) { og::set_error("%s: %s", __func__, msg); return OG_ERR_INVALID; } } while (0)

template <class T>
inline int upload(og_ctx* c, const char* name, const T* host, size_t count, T** dev)
{
    OG_TRY(c->reserve_t(name, std::max<size_t>(count, 1), dev));
    if (count) {
        OG_CUDA(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (int64_t)(count * sizeof(T));
    }
    return OG_OK;
}
template <class T>
inline int download(og_ctx* c, T* host, const T* dev, size_t count)
{
    if (count && host) {
        OG_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (int64_t)(count * sizeof(T));
    }
    return OG_OK;
}

static inline int h_nint(float x) { return (int)lroundf(x); }

// Radii actually used by the multi_scan loop of one slab (proc_oa.F90:641-707).
// Returns OG_ERR_RADIUS when the reference would STOP (surface scan 1 below 4.5 cells).
inline int scan_radii(const int* radius_influence, float sfc_mult, float pressure, int* out,
                      int* nscan)
{
    int rc = OG_OK, n = 0;
    for (int num_scan = 1; num_scan <= OG_MAX_SCAN; ++num_scan) {
        if (radius_influence[num_scan - 1] < 0) break;
        const bool is_sfc = fabsf(pressure - 1001.0f) < 0.0001f;
        volatile float sf = is_sfc ? sfc_mult : 1.00f;
        volatile float r_scaled = (float)radius_influence[num_scan - 1] * sf;
        volatile float r_scan1 = (float)radius_influence[0] * sf;
        if (sfc_mult < 0.99999f && is_sfc) {
            if (r_scan1 > 100.0f) { volatile float q = 100.0f / r_scan1; r_scaled = r_scaled * q; }
            if (r_scaled < 4.5f) { if (num_scan == 1) rc = OG_ERR_RADIUS; break; }
        }
        out[n++] = h_nint(r_scaled);
    }
    *nscan = n;
    return rc;
}

} // namespace og
