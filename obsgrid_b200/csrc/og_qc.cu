// og_qc.cu -- fused qc1 gross-error check + qc2 buddy check on sm_100a.
//
// Reference: src/qc1_module.F90:9-300 (error_max), src/qc2_module.F90:9-388 (buddy_check),
// src/qc0_module.F90:9-76 (flag arithmetic), :236-267 (local).
//
// The buddy means are FP32 sums over neighbours in ascending slab index; QC flags have to be
// bit-identical, so that order is kept: one thread owns one observation and walks, in index
// order through shared-memory tiles, the observations that can be within range of its cell --
// the slab is cut into cells one buddy range wide and og_bin.cu hands every cell the
// index-ordered list of the obs within range of it, so the reference's N^2 pair loop shrinks
// to N x (obs within ~3 ranges).  sqrt(dx*dx+dy*dy)*dx_km compared with the range and
// with 0.1 km is replaced by two thresholds on dx*dx+dy*dy found on the host by bisection
// over the float ordering with the very same operations, which is exact because
// sqrtf and the multiply are monotone.
//
// The reference relaxes difmax(num) by 1.2 inside the station loop when exactly two buddies
// survive (qc2_module.F90:327), and later stations read the relaxed value for buddies with a
// smaller index (:277-278).  That loop-carried dependency is resolved as a fix-point: the
// removal pass is re-run with the previous iteration's relaxation mask until the mask is
// stable (the dependency graph is a DAG ordered by index, so the fix-point is the
// sequential result).  Only buddies with diff > 1.25*difmax can ever be removed, so the
// removal pass sweeps a compacted, index-ordered list of those "suspects" only.
#include "og_ctx.h"
#include "og_device.cuh"

#include <algorithm>
#include <math.h>
#include <string.h>

namespace og {

struct QcSlabDev {
    const float* g;
    int n, ob_off, var;
    float pressure, max_error, max_buddy;
    float t_lo, t_hi;   // buddy iff t_lo <= dx*dx+dy*dy <= t_hi
    float reach;        // >= the largest |dx| or |dy| (grid cells) a buddy pair can have
    int ncx, ncy, cell_off;   // cells of the slab (edge ~ reach), first cell in the per-cell arrays
    int blk_off;              // first 256-ob block of the slab in the flat per-ob grids
    int cell;                 // cell edge in grid points
    int key_shift;            // home-list entry = (Morton code of the ob's place in its cell << key_shift) | index
};

struct QcDev {
    const QcSlabDev* slabs;
    int nslab;
    int iew, jns;
    float buddy_weight;
    int do_error_max, do_buddy;
    const float* xob;
    const float* yob;
    const float* val;
    const float* elev;
    const int* local_time;
    int* qc;
    float4* q4;          // {x, y, ob-aob, difmax0}
    int* emax_fail;      // error_max verdict per ob
    int* cnt;            // buddy_num
    float* sum;          // sum of buddy diffs (index order)
    unsigned char* susp; // ob can ever be removed from a buddy set
    int* susp_list;      // per cell (at range_start), local indices of its suspect candidates
    int* susp_count;     // per cell
    int4* items;         // work items {slab, cell, target group}
    int* n_items;
    unsigned char* mod_in;   // relaxation mask of the previous fix-point iteration
    unsigned char* mod_out;
    int* kob;            // buddy_num_final
    float* err;          // buddy err
    int* changed;
    int* home_key;       // entry of the ob in its cell's target list (see QcSlabDev::key_shift)
    short4* home_box;    // the grid cell an ob sits in
    short4* range_box;   // every grid cell within `reach` of it
    const int* home_count; const int* home_start; const int* home_list;      // targets of a cell
    const int* range_count; const int* range_start; const int* range_list;  // its candidates
    unsigned long long* pairs_examined;
    int sharded;         // og_set_exchange: only the obs of this rank's cells have cnt / kob / err
    float one;           // 1.0f the compiler cannot see (add2_of_product, og_device.cuh)
    // diagnostics (nullable)
    float* d_aob; float* d_err_max; float* d_thr_max; float* d_err_buddy; float* d_difmax;
    int* d_bnf;
};

__device__ __forceinline__ float dewpoint_factor(float ob)
{   // qc1_module.F90:156-166 == qc2_module.F90:177-186
    if ((ob < 255.0f) && (ob >= 235.0f)) return 1.25f;
    if ((ob < 235.0f) && (ob >= 215.0f)) return 1.50f;
    if (ob < 215.0f) return 1.75f;
    return 1.00f;
}

// error_max (qc1_module.F90:102-249) and the per-ob set-up of buddy_check (:121-213).
__global__ void __launch_bounds__(256) qc_prep_kernel(QcDev A)
{
    const QcSlabDev& S = A.slabs[slab_of_block(A.slabs, A.nslab, (int)blockIdx.x)];
    const int li = ((int)blockIdx.x - S.blk_off) * (int)blockDim.x + (int)threadIdx.x;
    if (li >= S.n) return;
    const int gi = S.ob_off + li;
    const float x = A.xob[gi], y = A.yob[gi], ob = A.val[gi];
    const float plevel = S.pressure;
    const float aob = bilinear(S.g, A.iew, x, y);
    const float err = fsub(ob, aob);

    int fail = 0;
    if (A.do_error_max) {
        float factor = 1.0f;
        switch (S.var) {
        case OG_VAR_UU: case OG_VAR_VV:
            if (plevel <= 499.9f) factor = 1.5f;
            else if (plevel <= 1000.1f) factor = 1.25f;
            break;
        case OG_VAR_TT:
            if (plevel > 1000.1f) {
                int lt = A.local_time[gi];
                factor = (lt >= 1100 && lt <= 2000) ? 1.5f : 1.25f;
            } else if ((plevel <= 1000.1f) && (plevel > 700.0f)) factor = 0.85f;
            else factor = 0.65f;
            break;
        case OG_VAR_PMSLPSFC:
            factor = fadd(1.0f, fdiv(fabsf(A.elev[gi]), 2000.0f));
            break;
        case OG_VAR_DEWPOINT:
            factor = dewpoint_factor(ob);
            break;
        default: break;
        }
        float thr = fmul(factor, S.max_error);
        if (S.var == OG_VAR_UU || S.var == OG_VAR_VV) thr = fmaxf(thr, fmul(0.15f, aob));
        fail = fabsf(err) > thr;
        if (A.d_err_max) A.d_err_max[gi] = err;
        if (A.d_thr_max) A.d_thr_max[gi] = thr;
    }
    A.emax_fail[gi] = fail;
    if (A.d_aob) A.d_aob[gi] = aob;

    float difmax = S.max_buddy;
    switch (S.var) { // qc2_module.F90:131-191
    case OG_VAR_UU: case OG_VAR_VV:
        if (plevel < 401.f) difmax = fmul(S.max_buddy, 1.7f);
        else if (plevel < 701.f) difmax = fmul(S.max_buddy, 1.4f);
        else if (plevel < 851.f) difmax = fmul(S.max_buddy, 1.1f);
        break;
    case OG_VAR_PMSLPSFC:
        difmax = fmul(S.max_buddy, fadd(1.0f, fdiv(A.elev[gi], 2000.0f)));
        break;
    case OG_VAR_TT:
        if (plevel < 401.f) difmax = fmul(S.max_buddy, 0.6f);
        else if (plevel < 701.f) difmax = fmul(S.max_buddy, 0.8f);
        break;
    case OG_VAR_DEWPOINT:
        difmax = fmul(S.max_buddy, dewpoint_factor(ob));
        break;
    default: break;
    }
    difmax = fmul(difmax, A.buddy_weight); // :193
    A.q4[gi] = make_float4(x, y, err, difmax);
    A.mod_in[gi] = 0;
    A.susp[gi] = (err > fmul(fminf(difmax, fmul(1.2f, difmax)), 1.25f)) ? 1 : 0;
    if (A.home_box) {
        // the same monotone clamp on both boxes keeps "within reach" => "lists meet" for any x, y
        const float fi = (float)A.iew, fj = (float)A.jns;
        const int hx = (int)fminf(fmaxf(floorf(x), 1.f), fi), hy = (int)fminf(fmaxf(floorf(y), 1.f), fj);
        A.home_box[gi] = make_short4((short)hx, (short)hx, (short)hy, (short)hy);
        // where in its cell: an 8 x 8 raster, Z-ordered, so that 32 consecutive targets of the sorted
        // list sit close together and their common bounding box culls the cell's candidates
        const int ox = (hx - 1) % S.cell, oy = (hy - 1) % S.cell;
        const unsigned sx = (unsigned)(ox * 8 / S.cell), sy = (unsigned)(oy * 8 / S.cell);
        const unsigned z = (sx & 1u) | ((sy & 1u) << 1) | ((sx & 2u) << 1) | ((sy & 2u) << 2) | ((sx & 4u) << 2) | ((sy & 4u) << 3);
        A.home_key[gi] = (int)(z << S.key_shift) | li;
        const float r = S.reach;
        A.range_box[gi] = make_short4((short)fminf(fmaxf(floorf(x - r) - 1.f, 1.f), fi),
                                      (short)fminf(fmaxf(ceilf(x + r) + 1.f, 1.f), fi),
                                      (short)fminf(fmaxf(floorf(y - r) - 1.f, 1.f), fj),
                                      (short)fminf(fmaxf(ceilf(y + r) + 1.f, 1.f), fj));
    }
}

// Work items of the two pair passes: (slab, cell, group of 32 targets of the cell), one WARP per item.
// Cells differ wildly in weight (targets x candidates grows with the square of the local station
// density): upper-air cells hold a dozen targets, a dense surface network hundreds -- with a warp as
// the unit neither leaves lanes idle, and the passes need no block barrier.
constexpr int BP_THREADS = 128;               // four items per CTA
constexpr int BP_WARPS = BP_THREADS / 32;
constexpr int BP_CHUNK = 128;                 // candidates staged per warp and round
__global__ void __launch_bounds__(128) buddy_items_kernel(QcDev A)
{
    const QcSlabDev& S = A.slabs[blockIdx.y];
    const int lc = blockIdx.x * blockDim.x + threadIdx.x;
    if (lc >= S.ncx * S.ncy) return;
    const int c = S.cell_off + lc;
    const int T = A.home_count[c];
    if (T == 0) return;
    const int G = (T + 31) / 32;
    const int base = atomicAdd(A.n_items, G);
    for (int g = 0; g < G; ++g) A.items[base + g] = make_int4((int)blockIdx.y, c, g, 0);
}

// station_loop_2/3, first half (qc2_module.F90:223-256): buddy count and index-ordered sum.
// Targets: 32 obs of a cell that sit close together (the cell's target list is Z-ordered in space);
// candidates: the index-ordered list of the obs within reach of the cell, culled -- in order -- to
// those within reach of the bounding box of the 32 targets.  The reference tests N - 1 candidates per
// target; a cell list holds (2.5 reach)^2 of area where the network is dense, the culled list about
// 3.8 reach^2, against the pi reach^2 that are buddies.
#ifndef OG_BP_STAGE
#define OG_BP_STAGE 128
#endif
constexpr int BP_STAGE = OG_BP_STAGE;     // candidates fetched per round (four per lane, all in flight)
__global__ void __launch_bounds__(BP_THREADS) buddy_pass1_kernel(QcDev A)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = (int)blockIdx.x * BP_WARPS + warp;
    if (item >= *A.n_items) return;
    const int4 it = A.items[item];
    const QcSlabDev& S = A.slabs[it.x];
    const int c = it.y, t0 = it.z * 32;
    const int T = A.home_count[c], L = A.range_count[c];
    const int* __restrict__ tl = A.home_list + A.home_start[c];
    const int* __restrict__ cl = A.range_list + A.range_start[c];
    const float tlo = S.t_lo, thi = S.t_hi;
    // the kept candidates as three arrays, so that the pair loop reads two candidates per load and
    // evaluates them with packed arithmetic against the target (the broadcast operand)
    constexpr int TCAP = BP_CHUNK + BP_STAGE;
    __shared__ __align__(16) float tile_all[BP_WARPS][3][TCAP];
    float* tx = tile_all[warp][0];
    float* ty = tile_all[warp][1];
    float* tz = tile_all[warp][2];
    const float one = A.one;
    const bool valid = t0 + lane < T;
    const int gi = S.ob_off + (valid ? (tl[t0 + lane] & ((1 << S.key_shift) - 1)) : 0);
    const float4 me = valid ? A.q4[gi] : make_float4(-1.e30f, -1.e30f, 0.f, 0.f);
    // bounding box of the group's targets
    float bx0 = valid ? me.x : 3.e38f, bx1 = valid ? me.x : -3.e38f, by0 = valid ? me.y : 3.e38f, by1 = valid ? me.y : -3.e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    // a candidate farther than this from the box is farther than t_hi from every target in it (the
    // margin swallows the roundings of both distance computations): it would add 0 to count and sum
    const float cull = thi * 1.0001f + 1.e-3f;
    const unsigned lt = (1u << lane) - 1u;
    const int nvalid = min(32, T - t0);
    int cnt = 0, fill = 0;
    float sum = 0.f;
    unsigned long long evaluated = 0;
    // (requesting the next round's candidates before this round's pairs are evaluated measured 2 %
    // slower: 74 registers instead of 68 cost more occupancy than the overlap returns)
    for (int base = 0; base < L; base += BP_STAGE) {
        float4 o[BP_STAGE / 32];
        bool keep[BP_STAGE / 32];
#pragma unroll
        for (int u = 0; u < BP_STAGE / 32; ++u) {
            const int k = base + 32 * u + lane;
            keep[u] = false;
            if (k < L) {
                o[u] = A.q4[S.ob_off + cl[k]];
                const float dxb = fmaxf(fmaxf(bx0 - o[u].x, o[u].x - bx1), 0.f), dyb = fmaxf(fmaxf(by0 - o[u].y, o[u].y - by1), 0.f);
                keep[u] = dxb * dxb + dyb * dyb <= cull;
            }
        }
#pragma unroll
        for (int u = 0; u < BP_STAGE / 32; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, keep[u]);
            if (keep[u]) { const int w = fill + __popc(m & lt); tx[w] = o[u].x; ty[w] = o[u].y; tz[w] = o[u].z; }
            fill += __popc(m);
        }
        if (fill > BP_CHUNK || base + BP_STAGE >= L) {
            const int lim8 = (fill + 7) & ~7;
            if (fill + lane < lim8) { tx[fill + lane] = 1.e30f; ty[fill + lane] = 1.e30f; tz[fill + lane] = 0.f; }
            __syncwarp();
            const f32x2 MX = bc(me.x), MY = bc(me.y);
#pragma unroll 4
            for (int q = 0; q < lim8; q += 2) {
                const f32x2 X = *reinterpret_cast<const f32x2*>(tx + q), Y = *reinterpret_cast<const f32x2*>(ty + q);
                const f32x2 DX = sub2(MX, X), DY = sub2(MY, Y);
                // dx*dx + dy*dy, each operation rounded as in qc2_module.F90:234-236
                float d0, d1, z0, z1;
                upk(add2_of_product(mul2(DX, DX), mul2(DY, DY), one), d0, d1);
                upk(*reinterpret_cast<const f32x2*>(tz + q), z0, z1);
                const bool b0 = (d0 <= thi) && (d0 >= tlo);   // self: d2 = 0 < t_lo
                const bool b1 = (d1 <= thi) && (d1 >= tlo);
                cnt += (b0 ? 1 : 0) + (b1 ? 1 : 0);
                sum = fadd(sum, b0 ? z0 : 0.f);
                sum = fadd(sum, b1 ? z1 : 0.f);
            }
            evaluated += (unsigned long long)fill;
            fill = 0;
            __syncwarp();
        }
    }
    if (valid) { A.cnt[gi] = cnt; A.sum[gi] = sum; }
    if (lane == 0) atomicAdd(A.pairs_examined, evaluated * (unsigned long long)nvalid);
}

// Per cell, the index-ordered sub-list of its candidates that can ever be removed from a
// buddy set: diff > 1.25*difmax with either the plain or the relaxed difmax
// (qc2_module.F90:277-282; flagged per ob by qc_prep_kernel).
__global__ void __launch_bounds__(128) suspect_kernel(QcDev A)
{
    const QcSlabDev& S = A.slabs[blockIdx.y];
    if ((int)blockIdx.x >= S.ncx * S.ncy) return;
    const int c = S.cell_off + blockIdx.x;
    const int L = (A.home_count[c] > 0) ? A.range_count[c] : 0;   // cells without targets are never read
    const int* __restrict__ cl = A.range_list + A.range_start[c];
    int* __restrict__ out = A.susp_list + A.range_start[c];
    __shared__ int s_w[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int done = 0;
    for (int base = 0; base < L; base += 128) {
        const int k = base + threadIdx.x;
        int lk = 0;
        bool sp = false;
        if (k < L) { lk = cl[k]; sp = A.susp[S.ob_off + lk] != 0; }
        const unsigned m = __ballot_sync(0xffffffffu, sp);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) { int v = s_w[w]; woff += (w < warp) ? v : 0; tot += v; }
        if (sp) out[done + woff + __popc(m & ((1u << lane) - 1u))] = lk;
        done += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) A.susp_count[c] = done;
}

// remove_bad_ob + err (qc2_module.F90:269-327) for one relaxation mask.
__global__ void __launch_bounds__(BP_THREADS) buddy_pass2_kernel(QcDev A)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = (int)blockIdx.x * BP_WARPS + warp;
    if (item >= *A.n_items) return;
    const int4 it = A.items[item];
    const QcSlabDev& S = A.slabs[it.x];
    const int c = it.y, t0 = it.z * 32;
    const int T = A.home_count[c];
    const bool valid = t0 + lane < T;
    const int li = valid ? (A.home_list[A.home_start[c] + t0 + lane] & ((1 << S.key_shift) - 1)) : 0;
    const int gi = S.ob_off + li;
    const float4 me = valid ? A.q4[gi] : make_float4(-1.e30f, -1.e30f, 0.f, 0.f);
    const int cnt = valid ? A.cnt[gi] : 0;
    float sum = valid ? A.sum[gi] : 0.f;
    const float average = (cnt > 0) ? fdiv(sum, (float)cnt) : 0.f;
    const float tlo = S.t_lo, thi = S.t_hi;
    const int ns = A.susp_count[c];
    const int* __restrict__ sl = A.susp_list + A.range_start[c];
    __shared__ float4 tile_all[BP_WARPS][32];
    __shared__ int tidx_all[BP_WARPS][32];   // local index, bit 30 = relaxed
    float4* tile = tile_all[warp];
    int* tidx = tidx_all[warp];
    int kob = cnt;
    for (int base = 0; base < ns; base += 32) {
        const int k = base + lane;
        __syncwarp();
        if (k < ns) {
            int lk = sl[k];
            tile[lane] = A.q4[S.ob_off + lk];
            tidx[lane] = lk | (A.mod_in[S.ob_off + lk] ? (1 << 30) : 0);
        }
        __syncwarp();
        const int lim = min(32, ns - base);
        if (cnt >= 2) {
            for (int q = 0; q < lim; ++q) {
                const float4 o = tile[q];
                const float dx = fsub(me.x, o.x), dy = fsub(me.y, o.y);
                const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
                if ((d2 <= thi) && (d2 >= tlo)) {
                    const int ti = tidx[q];
                    const int lk = ti & ~(1 << 30);
                    // stations before `num` have already been relaxed when num is visited
                    const float dm = ((ti >> 30) && lk < li) ? fmul(1.2f, o.w) : o.w;
                    const float diff_numj = fabsf(fsub(o.z, average));
                    if (o.z > fmul(dm, 1.25f) && diff_numj > dm) {
                        kob -= 1;
                        sum = fsub(sum, o.z);
                    }
                }
            }
        }
    }
    if (!valid) return;
    float err = 0.f;
    int bnf = 0;
    if (cnt >= 2) {
        bnf = kob;
        if (kob >= 2) err = fsub(me.z, fdiv(sum, (float)kob));
    }
    A.kob[gi] = bnf;
    A.err[gi] = err;
    const unsigned char m = (bnf == 2) ? 1 : 0;
    A.mod_out[gi] = m;
    if (!A.sharded && m != A.mod_in[gi]) *A.changed = 1;
}

// Sharded run: "did the relaxation mask move" over the combined mask, identically on every rank.
__global__ void __launch_bounds__(256) mask_moved_kernel(const unsigned char* __restrict__ a,
                                                         const unsigned char* __restrict__ b, int n, int* changed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool d = (i < n) && (a[i] != b[i]);
    if (__ballot_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) *changed = 1;
}

// Flags: error_max (2**16), no_buddies (2**14), buddy (2**17), in the reference's order.
__global__ void __launch_bounds__(256) qc_flag_kernel(QcDev A)
{
    const QcSlabDev& S = A.slabs[slab_of_block(A.slabs, A.nslab, (int)blockIdx.x)];
    const int li = ((int)blockIdx.x - S.blk_off) * (int)blockDim.x + (int)threadIdx.x;
    if (li >= S.n) return;
    const int gi = S.ob_off + li;
    int q = A.qc[gi];
    if (A.do_error_max && A.emax_fail[gi]) q = qc_add_flag(q, OG_FAILS_ERROR_MAX);
    if (A.do_buddy && !(A.sharded && A.kob[gi] == -1)) {   // -1: another rank's observation
        const int cnt = A.cnt[gi], bnf = A.kob[gi];
        const float err = A.err[gi];
        float dm = A.q4[gi].w;
        if (cnt < 2 || bnf < 2) q = qc_add_flag(q, OG_NO_BUDDIES);   // :309 / :322
        if (bnf == 2) dm = fmul(1.2f, dm);                           // :327
        if (fabsf(err) > dm) q = qc_add_flag(q, OG_FAILS_BUDDY_CHECK);
        if (A.d_err_buddy) A.d_err_buddy[gi] = err;
        if (A.d_difmax) A.d_difmax[gi] = dm;
        if (A.d_bnf) A.d_bnf[gi] = bnf;
    }
    A.qc[gi] = q;
}

// qc0_module.F90:261-267
__global__ void local_time_kernel(int hh, const float* __restrict__ lon, int n, int* __restrict__ out)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int lt = (int)fadd((float)hh, fdiv(lon[q], 15.f));
    if (lt < 0) lt = 24 + lt;
    out[q] = (lt % 24) * 100;
}

// GET_FG_T_AT_POINT, get_fg_at_point_module.F90:36-114 (after latlon_to_ij): one thread per point.
__global__ void fg_t_at_point_kernel(const float* __restrict__ t3, const float* __restrict__ h3,
                                     int jns, int iew, int kbu, const float* __restrict__ xs,
                                     const float* __restrict__ ys, const float* __restrict__ hs,
                                     int n, float* __restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float x = xs[q], y = ys[q], height = hs[q];
    const float dTdz = -0.0065f;            // constants.inc: dTdz_std_atm_lt_11000m
    if ((x < 1.f) || (x > (float)(iew - 1)) || (y < 1.f) || (y > (float)(jns - 1))) { out[q] = OG_MISSING_R; return; }
    const int fx = (int)floorf(x), cx = (int)ceilf(x), fy = (int)floorf(y), cy = (int)ceilf(y);
    const int nx[4] = {fx, fx, cx, cx}, ny[4] = {fy, cy, fy, cy};
    float tn[4];
    const size_t lev = (size_t)iew * jns;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const size_t col = (size_t)(nx[c] - 1) * jns + (size_t)(ny[c] - 1);   // (j,i,k): j fastest
        bool found = false;
        float hb = 0.f, ha = 0.f, tb = 0.f, ta = 0.f;
        for (int k = 2; k <= kbu - 1; ++k) {
            const float h0 = h3[col + (size_t)(k - 1) * lev], h1 = h3[col + (size_t)k * lev];
            if ((height >= h0) && (height <= h1)) {
                found = true; hb = h0; ha = h1;
                tb = t3[col + (size_t)(k - 1) * lev]; ta = t3[col + (size_t)k * lev];
                break;
            }
        }
        if (found) {
            const float w = fdiv(fsub(ha, height), fsub(ha, hb));
            tn[c] = fadd(fmul(w, tb), fmul(fsub(1.f, w), ta));
        } else if (height < h3[col + lev]) {
            tn[c] = fsub(t3[col + lev], fmul(dTdz, fsub(h3[col + lev], height)));
        } else {
            out[q] = OG_MISSING_R;          // above the top level (:93-96); the reference STOPs otherwise
            return;
        }
    }
    const float xn = fsub(x, floorf(x)), yn = fsub(y, floorf(y));
    out[q] = fadd(fmul(fsub(1.f, xn), fadd(fmul(fsub(1.f, yn), tn[0]), fmul(yn, tn[1]))),
                  fmul(xn, fadd(fmul(fsub(1.f, yn), tn[2]), fmul(yn, tn[3]))));
}

int fg_t_at_point_device(og_ctx* ctx, const float* t3, const float* h3, int jns, int iew, int kbu,
                         const float* x, const float* y, const float* h, int n, float* out)
{
    if (n <= 0) return OG_OK;
    fg_t_at_point_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(t3, h3, jns, iew, kbu, x, y, h, n, out);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

int local_time_device(og_ctx* ctx, int time_hhmm, const float* lon, int n, int* out)
{
    if (n <= 0) return OG_OK;
    local_time_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(time_hhmm / 100, lon, n, out);
    ctx->count_launch();
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

// ---- host: exact thresholds on d2 for  lo <= sqrtf(d2)*dx <= hi  ------------------------
static float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static float dist_of(float d2, float dx)
{
    volatile float s = sqrtf(d2);
    volatile float d = s * dx;
    return d;
}
// largest non-negative finite float d2 with dist_of(d2) <= lim (or -1 if none)
static float threshold_le(float lim, float dx)
{
    if (!(dist_of(0.f, dx) <= lim)) return -1.f;
    uint32_t lo = 0, hi = 0x7f7fffffu; // invariant: pred(lo) true
    if (dist_of(bits2f(hi), dx) <= lim) return bits2f(hi);
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (dist_of(bits2f(mid), dx) <= lim) lo = mid; else hi = mid;
    }
    return bits2f(lo);
}
// smallest non-negative float d2 with dist_of(d2) >= lim (or +inf if none)
static float threshold_ge(float lim, float dx)
{
    if (dist_of(0.f, dx) >= lim) return 0.f;
    uint32_t lo = 0, hi = 0x7f7fffffu; // invariant: pred(lo) false
    if (!(dist_of(bits2f(hi), dx) >= lim)) return INFINITY;
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (dist_of(bits2f(mid), dx) >= lim) hi = mid; else lo = mid;
    }
    return bits2f(hi);
}

int qc_run_batch(og_ctx* ctx, QcBatch& b)
{
    const int ns = (int)b.slabs.size();
    if (ns == 0) return OG_OK;
    // cell boxes are packed into short4, slabs are gridDim.y of the per-cell kernels
    if (b.iew > 32000 || b.jns > 32000) { set_error("grid larger than 32000 points per edge"); return OG_ERR_INVALID; }
    if (ns > 65535) { set_error("more than 65535 slabs in one batch"); return OG_ERR_INVALID; }
    cudaStream_t st = ctx->stream;
    std::vector<QcSlabDev> hs(ns);
    int max_n = 0, ncells_all = 0, max_cells = 0, nblk_ob = 0;
    std::vector<BinSlab> bslabs(ns);
    int64_t pair_tests = 0, nobs = 0;
    // thresholds depend on the range class only: cache the three
    float thi_c[3], tlo = threshold_ge(0.1f, b.dx_km);
    const float ranges[3] = {650.f, 300.f, 500.f};
    for (int r = 0; r < 3; ++r) thi_c[r] = threshold_le(ranges[r], b.dx_km);
    for (int s = 0; s < ns; ++s) {
        const QcSlabHost& h = b.slabs[s];
        QcSlabDev& d = hs[s];
        d.g = h.g; d.n = h.n; d.ob_off = h.ob_off; d.var = h.var; d.pressure = h.pressure;
        d.max_error = h.max_error; d.max_buddy = h.max_buddy;
        int rc = (h.pressure <= 201.0f) ? 0 : ((h.pressure <= 1000.1f) ? 1 : 2); // :121-127
        d.t_lo = tlo; d.t_hi = thi_c[rc];
        // |dx| <= sqrt(d2) for a pair that passes d2 <= t_hi; the margin swallows the roundings
        const float reach_max = (float)std::max(b.iew, b.jns) + 2.f;
        d.reach = (d.t_hi >= 0.f) ? std::min(sqrtf(d.t_hi) * 1.0001f + 0.01f, reach_max) : 0.f;
        // cells one reach wide (9 cells per ob) where stations are sparse, half a reach wide
        // (25 cells per ob, a third fewer pair tests) where a half-reach cell still fills a warp
        // cells one reach wide (9 cells per ob) where stations are sparse, half a reach wide (25 cells
        // per ob, a third fewer pair tests) where a half-reach cell still fills a warp.  Finer cells do
        // not pay: a third / a quarter of a reach examine 11 % / 16 % fewer pairs (cfg4: pass 1 26.6 ->
        // 24.9 ms) but the index-ordered lists grow 2x / 3.2x (QC stage 30.5 -> 31.2 / 34.3 ms;
        // profiles/exp/celldiv.sh)
        const float per_reach2 = (float)h.n / ((float)b.iew * (float)b.jns) * d.reach * d.reach;
        const int div = (per_reach2 / 4.f >= 32.f) ? 2 : 1;
        const int cell = std::max(8, (int)ceilf(d.reach / (float)div));
        d.ncx = (b.iew + cell - 1) / cell; d.ncy = (b.jns + cell - 1) / cell;
        d.cell = cell;
        d.key_shift = 1;
        while ((1 << d.key_shift) < std::max(h.n, 2)) ++d.key_shift;
        if (d.key_shift > 25) { set_error("more than 2^25 observations in one buddy-check slab"); return OG_ERR_INVALID; }
        d.cell_off = ncells_all;
        d.blk_off = nblk_ob;
        nblk_ob += (h.n + 255) / 256;
        bslabs[s] = BinSlab{h.n, h.ob_off, cell, d.ncx, d.ncy, d.cell_off};
        ncells_all += d.ncx * d.ncy;
        max_cells = std::max(max_cells, d.ncx * d.ncy);
        max_n = std::max(max_n, h.n);
        pair_tests += (int64_t)h.n * (h.n > 0 ? h.n - 1 : 0);
        nobs += h.n;
    }
    if (max_n == 0) return OG_OK;
    const size_t tot = (size_t)std::max(b.total_obs, 1);
    QcSlabDev* d_slabs = nullptr;
    OG_TRY(ctx->reserve_t("qc.slabs", (size_t)ns, &d_slabs));
    QcDev A{};
    A.slabs = d_slabs; A.nslab = ns; A.iew = b.iew; A.jns = b.jns; A.buddy_weight = b.buddy_weight;
    A.do_error_max = b.do_error_max; A.do_buddy = b.do_buddy;
    A.xob = b.xob; A.yob = b.yob; A.val = b.val; A.elev = b.elev; A.local_time = b.local_time;
    A.qc = b.qc;
    OG_TRY(ctx->reserve_t("qc.q4", tot, &A.q4));
    OG_TRY(ctx->reserve_t("qc.emax", tot, &A.emax_fail));
    OG_TRY(ctx->reserve_t("qc.cnt", tot, &A.cnt));
    OG_TRY(ctx->reserve_t("qc.sum", tot, &A.sum));
    OG_TRY(ctx->reserve_t("qc.suspflag", tot, &A.susp));
    OG_TRY(ctx->reserve_t("qc.mod0", tot, &A.mod_in));
    OG_TRY(ctx->reserve_t("qc.mod1", tot, &A.mod_out));
    OG_TRY(ctx->reserve_t("qc.kob", tot, &A.kob));
    OG_TRY(ctx->reserve_t("qc.err", tot, &A.err));
    A.changed = ctx->d_flag;
    if (b.do_buddy) {
        OG_TRY(ctx->reserve_t("qc.home_box", tot, &A.home_box));
        OG_TRY(ctx->reserve_t("qc.home_key", tot, &A.home_key));
        OG_TRY(ctx->reserve_t("qc.range_box", tot, &A.range_box));
    }
    A.pairs_examined = ctx->d_counters + 3;
    const bool sharded = ctx->shard_world > 1 && ctx->exchange && b.do_buddy;
    A.sharded = sharded ? 1 : 0;
    A.one = 1.0f;
    auto exchange = [&](void* p, size_t n, int dtype, int op) -> int {
        if (ctx->exchange(ctx->exchange_user, p, n, dtype, op, (void*)st) != 0) {
            set_error("exchange callback failed"); return OG_ERR_CUDA;
        }
        return OG_OK;
    };
    A.d_aob = b.aob; A.d_err_max = b.err_max; A.d_thr_max = b.thr_max;
    A.d_err_buddy = b.err_buddy; A.d_difmax = b.difmax; A.d_bnf = b.buddy_num_final;

    OG_TRY(ctx->put_small(d_slabs, hs.data(), sizeof(QcSlabDev) * ns));
    qc_prep_kernel<<<nblk_ob, 256, 0, st>>>(A);
    ctx->count_launch();
    if (b.do_buddy) {
        // targets of a cell: the obs sitting in it (any order); candidates: the obs within reach
        // of it, in ascending index -- the order the reference sums in
        BinLists home, range;
        // sharded: this rank's cells only, so the targets (and their candidate lists) are dealt out
        const int orank = sharded ? ctx->shard_rank : 0, oworld = sharded ? ctx->shard_world : 1;
        OG_TRY(bin_build(ctx, "qc.home", bslabs, A.home_box, true, &home, orank, oworld, A.home_key));
        OG_TRY(bin_build(ctx, "qc.range", bslabs, A.range_box, true, &range, orank, oworld));
        if (sharded) OG_CUDA(cudaMemsetAsync(A.kob, 0xff, tot * sizeof(int), st));
        A.home_count = home.cell_count; A.home_start = home.cell_start; A.home_list = home.cell_list;
        A.range_count = range.cell_count; A.range_start = range.cell_start; A.range_list = range.cell_list;
        OG_TRY(ctx->reserve_t("qc.susp_list", (size_t)std::max(range.list_total, 1), &A.susp_list));
        OG_TRY(ctx->reserve_t("qc.susp_count", (size_t)ncells_all + 1, &A.susp_count));
        // one work item per (cell, group of 32 targets): at most one per non-empty cell plus one per
        // 32 obs; BP_WARPS items per CTA
        const size_t max_items = (size_t)ncells_all + (size_t)b.total_obs / 32 + 1;
        const unsigned pair_grid = (unsigned)((max_items + BP_WARPS - 1) / BP_WARPS);
        OG_TRY(ctx->reserve_t("qc.items", max_items, &A.items));
        A.n_items = ctx->d_flag + 2;
        OG_CUDA(cudaMemsetAsync(A.n_items, 0, sizeof(int), st));
        dim3 g_ci((max_cells + 127) / 128, ns);
        buddy_items_kernel<<<g_ci, 128, 0, st>>>(A);
        const int tslot = ctx->timer_begin(1);
        buddy_pass1_kernel<<<pair_grid, BP_THREADS, 0, st>>>(A);
        ctx->timer_end(tslot);
        suspect_kernel<<<dim3(max_cells, ns), 128, 0, st>>>(A);
        ctx->count_launch(3);
        int iters = 0;
        for (;;) {
            OG_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st));
            if (sharded) OG_CUDA(cudaMemsetAsync(A.mod_out, 0, tot, st));
            buddy_pass2_kernel<<<pair_grid, BP_THREADS, 0, st>>>(A);
            ctx->count_launch();
            if (sharded) {   // qc2_module.F90:327's mask of every observation, whoever owns it
                OG_TRY(exchange(A.mod_out, tot, OG_XCHG_UINT8, OG_XCHG_MAX));
                mask_moved_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(A.mod_in, A.mod_out, (int)tot, A.changed);
                ctx->count_launch();
            }
            void* h_changed = nullptr;
            OG_TRY(ctx->get_small(ctx->d_flag, sizeof(int), &h_changed));
            OG_CUDA(cudaStreamSynchronize(st));
            ++iters;
            if (!*(const int*)h_changed) break;
            std::swap(A.mod_in, A.mod_out);
            if (iters > max_n + 2) { set_error("buddy difmax fix-point did not converge"); return OG_ERR_NOCONV; }
        }
        ctx->stats.buddy_fixpoint_iters += iters;
        ctx->stats.buddy_pair_tests += pair_tests;
        ctx->stats.buddy_obs += nobs;
    }
    qc_flag_kernel<<<nblk_ob, 256, 0, st>>>(A);
    ctx->count_launch();
    // every rank started from the same flags and only the owner of an observation added buddy bits
    if (sharded) OG_TRY(exchange(b.qc, tot, OG_XCHG_INT32, OG_XCHG_MAX));
    OG_CUDA(cudaGetLastError());
    return OG_OK;
}

} // namespace og
