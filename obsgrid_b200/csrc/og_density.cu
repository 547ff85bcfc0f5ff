// og_density.cu -- observation density and distance to the nearest observed grid box for the
// surface FDDA: ob_density / obs_distance, src/qc0_module.F90:81-232.
//
// ob_density: the reference sweeps a window per ob and increments tobbox where
// SQRT((i-x)**2 + (j-y)**2) < 250/dx.  Here a grid point counts the obs of its cell list
// (og_bin.cu, boxes = the reference's windows); counts are integers held exactly in FP32, so
// the order is irrelevant.  The window (INT(x) +- (NINT(r)+1)) always contains the disc, so the
// only clipping is the domain's (1 <= i <= iew-1, 1 <= j <= jns-1).
//
// obs_distance: the reference is O((iew*jns)^2): for every empty box the minimum over all
// observed boxes of SQRT(REAL(dj)**2 + REAL(di)**2), times dx.  Squared integer distances are
// exact, REAL(dj)**2 + REAL(di)**2 is RN of their sum and SQRT / MIN are monotone, so the
// result is sqrtf(RN(D2min)) * dx with D2min from an exact two-pass Euclidean distance
// transform: per row the distance along i to the nearest observed box, then per column the
// lower envelope min_jj (g(jj)^2 + (jj-j)^2), walked outwards from j until jj's own offset
// exceeds the best found.
#include "og_ctx.h"
#include "og_device.cuh"

#include <algorithm>
#include <math.h>

namespace og {

namespace {

constexpr int DCELL = 32;
constexpr int BIGD2 = 1 << 30;       // "no observed box in this row"

struct DensDev {
    const float* x; const float* y;
    int n, iew, jns, ncx, ncy;
    float r;                 // 250. / grid_dist
    short4* box;
    const int* cell_count; const int* cell_start; const int* cell_list;
    float* tobbox;           // (jns, iew), j fastest, in/out
};

__global__ void __launch_bounds__(256) density_box_kernel(DensDev A, int rr)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.n) return;
    const int iob = (int)A.x[q], job = (int)A.y[q];
    const int i0 = max(iob - rr - 1, 1), i1 = min(iob + rr + 1, A.iew - 1);
    const int j0 = max(job - rr - 1, 1), j1 = min(job + rr + 1, A.jns - 1);
    A.box[q] = make_short4((short)max(i0, -32000), (short)min(i1, 32000), (short)max(j0, -32000), (short)min(j1, 32000));
}

// One CTA per 32x32 cell; thread t owns 4 points of a column segment so that the tobbox
// accesses (j fastest) are coalesced along j.
__global__ void __launch_bounds__(256) density_count_kernel(DensDev A)
{
    const int c = blockIdx.x;
    const int cx = c % A.ncx, cy = c / A.ncx;
    const int L = A.cell_count[c];
    if (L == 0) return;
    const int* __restrict__ list = A.cell_list + A.cell_start[c];
    __shared__ float2 s_xy[256];
    const int j = cy * DCELL + (threadIdx.x & 31) + 1;
    const int ib = cx * DCELL + (threadIdx.x >> 5) * 4 + 1;          // 4 consecutive i
    const float fj = (float)j;
    int cnt[4] = {0, 0, 0, 0};
    for (int base = 0; base < L; base += 256) {
        __syncthreads();
        const int k = base + threadIdx.x;
        if (k < L) { const int q = list[k]; s_xy[threadIdx.x] = make_float2(A.x[q], A.y[q]); }
        __syncthreads();
        const int lim = min(256, L - base);
        for (int t = 0; t < lim; ++t) {
            const float2 o = s_xy[t];
            const float dy = fsub(fj, o.y), dy2 = fmul(dy, dy);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = fsub((float)(ib + u), o.x);
                cnt[u] += (fsqrt(fadd(fmul(dx, dx), dy2)) < A.r) ? 1 : 0;
            }
        }
    }
    if (j <= A.jns - 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = ib + u;
            if (i <= A.iew - 1 && cnt[u]) {
                float* p = A.tobbox + (size_t)(i - 1) * A.jns + (size_t)(j - 1);
                *p = fadd(*p, (float)cnt[u]);     // == cnt[u] increments by one: integers < 2^24
            }
        }
    }
}

// Pass 1: per row j, squared distance along i to the nearest observed box (BIGD2 if none).
// One thread per j: neighbouring threads touch neighbouring addresses (j fastest).
__global__ void __launch_bounds__(128) edt_rows_kernel(const float* __restrict__ tobbox, int* __restrict__ g2,
                                                       int iew, int jns)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= jns) return;
    int last = -1;
    for (int i = 0; i < iew; ++i) {               // forward: nearest observed box at or before i
        const size_t o = (size_t)i * jns + j;
        if (tobbox[o] > 0.5f) last = i;
        g2[o] = (last < 0) ? BIGD2 : (i - last) * (i - last);
    }
    last = -1;
    for (int i = iew - 1; i >= 0; --i) {          // backward: at or after i
        const size_t o = (size_t)i * jns + j;
        if (tobbox[o] > 0.5f) last = i;
        if (last >= 0) g2[o] = min(g2[o], (last - i) * (last - i));
    }
}

// Pass 2: per column i (contiguous in j), D2(j) = min_jj g2(jj) + (jj-j)^2; distance = RN-sqrt * dx.
__global__ void __launch_bounds__(256) edt_cols_kernel(const int* __restrict__ g2, float* __restrict__ distance,
                                                       int iew, int jns, float dx, float diag_cells)
{
    extern __shared__ int s_g[];
    const int i = blockIdx.x;
    const int* __restrict__ col = g2 + (size_t)i * jns;
    for (int j = threadIdx.x; j < jns; j += blockDim.x) s_g[j] = col[j];
    __syncthreads();
    for (int j = threadIdx.x; j < jns; j += blockDim.x) {
        int best = s_g[j];
        for (int d = 1; d < jns; ++d) {
            const int d2 = d * d;
            if (d2 >= best) break;
            if (j - d >= 0) best = min(best, s_g[j - d] + d2);
            if (j + d < jns) best = min(best, s_g[j + d] + d2);
        }
        float out;
        if (best == 0) out = 0.0f;
        else {
            float dijmin = diag_cells;
            if (best < BIGD2) dijmin = fminf(dijmin, fsqrt((float)best));
            out = fmul(dijmin, dx);
        }
        distance[(size_t)i * jns + j] = out;
    }
}

} // namespace

int ob_density_device(og_ctx* ctx, const float* d_x, const float* d_y, float grid_dist, int n,
                      float* d_tobbox, int iew, int jns)
{
    if (n <= 0) return OG_OK;
    DensDev A{};
    A.x = d_x; A.y = d_y; A.n = n; A.iew = iew; A.jns = jns;
    A.ncx = (iew + DCELL - 1) / DCELL; A.ncy = (jns + DCELL - 1) / DCELL;
    volatile float r = 250.f / grid_dist;
    A.r = r;
    const int rr = (int)lroundf(r);
    A.tobbox = d_tobbox;
    OG_TRY(ctx->reserve_t("dens.box", (size_t)n, &A.box));
    density_box_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(A, rr);
    ctx->count_launch();
    std::vector<BinSlab> bs(1, BinSlab{n, 0, DCELL, A.ncx, A.ncy, 0});
    BinLists bl;
    OG_TRY(bin_build(ctx, "dens.bin", bs, A.box, false, &bl));
    A.cell_count = bl.cell_count; A.cell_start = bl.cell_start; A.cell_list = bl.cell_list;
    density_count_kernel<<<A.ncx * A.ncy, 256, 0, ctx->stream>>>(A);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

int obs_distance_device(og_ctx* ctx, const float* d_tobbox, float* d_distance, int iew, int jns, float dx)
{
    // sqrtf((float)(dj*dj + di*di)) equals the reference's SQRT(REAL(dj)*REAL(dj) + REAL(di)*REAL(di))
    // only while each square is exact in FP32 (< 2**24): the bit-identical guarantee ends at 4096
    // points per edge -- BASELINE's largest grid is 4000 x 4000 -- and larger grids are refused
    if (iew > 4096 || jns > 4096) { set_error("obs_distance: grids beyond 4096 points per edge are not supported"); return OG_ERR_INVALID; }
    int* g2 = nullptr;
    OG_TRY(ctx->reserve_t("dens.g2", (size_t)iew * jns, &g2));
    // domain_diagonal_m / dx with the reference's evaluation order (qc0_module.F90:151-152)
    volatile float a = (float)(iew - 1) * dx; a = a * (float)(iew - 1); a = a * dx;
    volatile float b = (float)(jns - 1) * dx; b = b * (float)(jns - 1); b = b * dx;
    volatile float s = a + b;
    volatile float diag_m = sqrtf(s);
    volatile float diag_cells = diag_m / dx;
    edt_rows_kernel<<<(jns + 127) / 128, 128, 0, ctx->stream>>>(d_tobbox, g2, iew, jns);
    const size_t smem = sizeof(int) * (size_t)jns;
    if (smem > 48 * 1024)
        OG_CUDA(cudaFuncSetAttribute(edt_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edt_cols_kernel<<<iew, 256, smem, ctx->stream>>>(g2, d_distance, iew, jns, dx, diag_cells);
    ctx->count_launch(2);
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

} // namespace og
