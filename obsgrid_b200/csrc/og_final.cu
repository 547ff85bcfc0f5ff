// og_final.cu -- post-analysis epilogue on the device: the per-grid-point maps
// that sit between proc_oa and the netCDF writer.
//
// Reference: src/final_analysis_module.F90:316-350 with src/first_guess_module.F90:59-91 (A-grid
// wind increments averaged onto the C-grid winds), :364-370 (surface level of PRES follows the SLP
// increment), :1026-1201 (mixing_ratio + sfcprs).  3-D arrays keep the reference layout
// (jns,iew,kbu), j fastest; threads run along j, so every access is coalesced.
//
// Arithmetic: the wind and pressure maps are FP32 round-to-nearest in source order (bit-identical
// to the oracle).  mixing_ratio / sfcprs call EXP, LOG and ** -- glibc's expf/logf/powf in the
// gfortran build -- which are evaluated here in FP64 and rounded once.
#include "og_ctx.h"
#include "og_device.cuh"

namespace og {

namespace {

__device__ __forceinline__ size_t ix3(int j, int i, int k, int jns, int iew)
{
    return ((size_t)k * (size_t)iew + (size_t)i) * (size_t)jns + (size_t)j;
}

// DIR 0: crs2dot_pertU, average along dim2 (iew); DIR 1: crs2dot_pertV, along dim1 (jns).
template <int DIR>
__global__ void __launch_bounds__(256) wind_increment_kernel(const float* __restrict__ w,
                                                             const float* __restrict__ wA,
                                                             const float* __restrict__ wC,
                                                             float* __restrict__ out, int jns, int iew)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, k = blockIdx.z;
    if (j >= jns) return;
    const int n = DIR ? jns : iew, p = DIR ? j : i;
    auto pert = [&](int q) {
        const size_t o = DIR ? ix3(q, i, k, jns, iew) : ix3(j, q, k, jns, iew);
        return fsub(w[o], wA[o]);
    };
    float v;
    if (p == 0) v = pert(0);
    else if (p == n - 1) v = pert(n - 2);
    else v = fmul(fadd(pert(p - 1), pert(p)), 0.5f);
    const size_t o = ix3(j, i, k, jns, iew);
    out[o] = fadd(wC[o], v);
}

__global__ void pres_sfc_kernel(const float* __restrict__ pres, const float* __restrict__ slp_C,
                                const float* __restrict__ slp_x, float* __restrict__ out, size_t n)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[q] = fadd(pres[q], fmul(fdiv(pres[q], slp_C[q]), fsub(slp_x[q], slp_C[q])));
}

// temporal_interp (tinterp_module.F90:126-213): (w1*a + w2*b)/wd; cross-point fields are computed on
// (1:jns-1, 1:iew-1) and their last row / column replicated (:131-132), i.e. read from the
// clamped source point.
__global__ void __launch_bounds__(256) temporal_interp_kernel(const float* __restrict__ a,
                                                              const float* __restrict__ b,
                                                              float* __restrict__ out, int jns, int iew,
                                                              float w1, float w2, float wd, int cross)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, k = blockIdx.z;
    if (j >= jns) return;
    const int js = cross ? min(j, jns - 2) : j, is = cross ? min(i, iew - 2) : i;
    const size_t o = ix3(js, is, k, jns, iew);
    out[ix3(j, i, k, jns, iew)] = fdiv(fadd(fmul(w1, a[o]), fmul(w2, b[o])), wd);
}

__device__ __forceinline__ float exp_rn(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_rn(float x) { return (float)log((double)x); }
__device__ __forceinline__ float pow_rn(float x, float y) { return (float)pow((double)x, (double)y); }

__device__ __forceinline__ float sat_vapor_pressure(float t)
{
    const float svp1x10 = fmul(0.6112f, 10.f), svp2 = 17.67f, svp3 = 29.65f, svpt0 = 273.15f;
    return fmul(svp1x10, exp_rn(fdiv(fmul(svp2, fsub(t, svpt0)), fsub(t, svp3))));
}

// One thread per cross point (j,i): rh clamp and qv of every level, sfcprs, qv of the surface.
__global__ void __launch_bounds__(128) mixing_ratio_kernel(float* __restrict__ rh,
                                                           const float* __restrict__ t,
                                                           const float* __restrict__ h,
                                                           const float* __restrict__ terrain,
                                                           const float* __restrict__ slp_x,
                                                           const float* __restrict__ pressure,
                                                           int iew, int jns, int kbu, int k850,
                                                           int k700, int k500, float p78, float p57,
                                                           float* __restrict__ qv,
                                                           float* __restrict__ psfc)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= jns - 1 || i >= iew - 1) return;
    const float eps = 0.622f;
    float q850 = 0.f, q700 = 0.f, q500 = 0.f, rh0 = 0.f;
    for (int k = 0; k < kbu; ++k) {
        const size_t o = ix3(j, i, k, jns, iew);
        const float r = fminf(fmaxf(rh[o], 1.f), 100.f);       // :1055
        rh[o] = r;
        if (k == 0) { rh0 = r; continue; }
        const float es = sat_vapor_pressure(t[o]);              // :1058-1060
        const float qs = fdiv(fmul(eps, es), fsub(pressure[k], es));
        const float q = fmaxf(fmul(fmul(0.01f, r), qs), 0.0f);
        qv[o] = q;
        if (k == k850) q850 = q;
        if (k == k700) q700 = q;
        if (k == k500) q500 = q;
    }
    // sfcprs, :1129-1199
    const float R = 287.f, g = 9.806f, gamma = 6.5e-3f;
    const float gammarg = fdiv(fmul(gamma, R), g);
    const size_t o2 = (size_t)i * jns + j;
    const float ter = terrain[o2], slp = slp_x[o2];
    const float ht = fdiv(-ter, h[ix3(j, i, k850, jns, iew)]);
    const float ps0 = fmul(slp, pow_rn(fdiv(slp, 85000.f), ht));
    float p1;
    if (ps0 > 95000.f) p1 = 85000.f;
    else if (ps0 > 70000.f) p1 = fsub(ps0, 10000.f);
    else p1 = 50000.f;
    const float t850 = fmul(t[ix3(j, i, k850, jns, iew)], fadd(1.f, fmul(0.608f, q850)));
    const float t700 = fmul(t[ix3(j, i, k700, jns, iew)], fadd(1.f, fmul(0.608f, q700)));
    const float t500 = fmul(t[ix3(j, i, k500, jns, iew)], fadd(1.f, fmul(0.608f, q500)));
    const float gamma78 = fmul(log_rn(fdiv(t850, t700)), p78);
    const float gamma57 = fmul(log_rn(fdiv(t700, t500)), p57);
    float t1;
    if (ps0 > 95000.f) t1 = t850;
    else if (ps0 > 85000.f) t1 = fmul(t700, pow_rn(fdiv(p1, 70000.f), gamma78));
    else if (ps0 > 70000.f) t1 = fmul(t500, pow_rn(fdiv(p1, 50000.f), gamma57));
    else t1 = t500;
    const float tslv = fmul(t1, pow_rn(fdiv(slp, p1), gammarg));
    const float tsfc = fsub(tslv, fmul(gamma, ter));
    const float ps = fmul(slp, exp_rn(fdiv(fmul(-ter, g), fmul(fdiv(R, 2.f), fadd(tsfc, tslv)))));
    psfc[o2] = ps;
    // surface qv, :1074-1080
    const size_t o = ix3(j, i, 0, jns, iew);
    const float es = sat_vapor_pressure(t[o]);
    const float qs = fdiv(fmul(eps, es), fsub(fdiv(ps, 100.f), es));
    qv[o] = fmaxf(fmul(fmul(0.01f, rh0), qs), 0.0f);
}

} // namespace

int wind_increment_device(og_ctx* ctx, int dir, const float* w, const float* wA, const float* wC,
                          float* out, int jns, int iew, int kbu)
{
    dim3 grid((jns + 255) / 256, iew, kbu);
    if (dir == 0) wind_increment_kernel<0><<<grid, 256, 0, ctx->stream>>>(w, wA, wC, out, jns, iew);
    else wind_increment_kernel<1><<<grid, 256, 0, ctx->stream>>>(w, wA, wC, out, jns, iew);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

int temporal_interp_device(og_ctx* ctx, const float* a, const float* b, float* out, int jns, int iew,
                           int nlev, float w1, float w2, float wd, int cross)
{
    dim3 grid((jns + 255) / 256, iew, nlev);
    temporal_interp_kernel<<<grid, 256, 0, ctx->stream>>>(a, b, out, jns, iew, w1, w2, wd, cross);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

int pres_sfc_device(og_ctx* ctx, const float* pres, const float* slp_C, const float* slp_x,
                    float* out, size_t n)
{
    pres_sfc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(pres, slp_C, slp_x, out, n);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

int mixing_ratio_device(og_ctx* ctx, float* rh, const float* t, const float* h, const float* terrain,
                        const float* slp_x, const float* d_pressure, const float* h_pressure, int iew,
                        int jns, int kbu, float* qv, float* psfc)
{
    // sfcprs :1108-1127: the 850, 700 and 500 hPa levels must exist (the reference STOPs)
    int k850 = -1, k700 = -1, k500 = -1;
    for (int k = 1; k < kbu; ++k) {
        const long p = lroundf(h_pressure[k]);
        if (p == 850) k850 = k; else if (p == 700) k700 = k; else if (p == 500) k500 = k;
    }
    if (k850 < 0 || k700 < 0 || k500 < 0) {
        set_error("mixing_ratio: no 850, 700 or 500 hPa level (final_analysis_module.F90:1122)");
        return OG_ERR_INVALID;
    }
    // p78 / p57 are compile-time constants in the reference (LOG of a literal quotient)
    const float p78 = 1.f / (float)log((double)(850.f / 700.f));
    const float p57 = 1.f / (float)log((double)(700.f / 500.f));
    dim3 grid((jns - 1 + 127) / 128, iew - 1);
    mixing_ratio_kernel<<<grid, 128, 0, ctx->stream>>>(rh, t, h, terrain, slp_x, d_pressure, iew, jns,
                                                      kbu, k850, k700, k500, p78, p57, qv, psfc);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

} // namespace og
