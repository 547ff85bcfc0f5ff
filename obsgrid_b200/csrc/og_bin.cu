// og_bin.cu -- stable spatial binning of observations on the device.
//
// Both hot loops of the reference sum over observations in ascending slab index
// (oa_module.F90:116, qc2_module.F90:230,250-252) and the results have to be bit-identical,
// so every spatial candidate list handed to a kernel must keep that order.  Given one
// bounding box per ob (1-based inclusive grid coordinates) and a cell size per slab, this
// builds, for every cell of every slab of a batch, the list of obs whose box touches the cell:
//   count (atomics) -> exclusive prefix over all cells of the batch -> fill (atomics, any
//   order) -> per-cell sort by ob index (shared-memory rank / bitonic sort).
#include "og_ctx.h"
#include "og_device.cuh"

#include <algorithm>
#include <string>

namespace og {

namespace {

constexpr int BIN_THREADS = 256;
constexpr int SORT_CAP = 8192;       // cell lists up to this length are sorted in shared memory
constexpr int NB_MAX = 4096;         // value buckets of the long-list sorter

struct BinDev {
    const BinSlab* slabs;
    int ns;
    int own_rank, own_world;   // cells dealt to ranks: (local cell index) % own_world == own_rank are built
    const short4* box;
    const int* key;     // nullable: list entry of an ob instead of its local index
    int* cell_count;
    int* cell_cursor;
    int* cell_start;
    int* cell_list;
    int* cell_sorted;
    int* big_cells;     // cells whose list is too long for the warp sorter
    int* n_big;
};

template <class F>
__device__ __forceinline__ void for_each_cell(const BinDev& A, const BinSlab& S, short4 b, F f)
{
    if (b.x > b.y || b.z > b.w) return;
    const int cx0 = max((b.x - 1) / S.cell, 0), cx1 = min((b.y - 1) / S.cell, S.ncx - 1);
    const int cy0 = max((b.z - 1) / S.cell, 0), cy1 = min((b.w - 1) / S.cell, S.ncy - 1);
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            const int lc = cy * S.ncx + cx;
            if (A.own_world > 1 && lc % A.own_world != A.own_rank) continue;
            f(S.cell_off + lc);
        }
}

// Slabs with at most HCAP cells histogram in shared memory first, so that the global atomics
// are one per (block, touched cell) instead of one per (ob, cell): with cells one buddy range
// wide a few hundred counters would otherwise take every ob of the slab nine times.
constexpr int HCAP = 4096;

__global__ void __launch_bounds__(256) bin_count_kernel(BinDev A)
{
    const BinSlab S = A.slabs[slab_of_block(A.slabs, A.ns, (int)blockIdx.x)];
    const int li = ((int)blockIdx.x - S.blk_off) * (int)blockDim.x + (int)threadIdx.x;
    const int nc = S.ncx * S.ncy;
    const short4 box = (li < S.n) ? A.box[S.ob_off + li] : make_short4(1, 0, 1, 0);
    if (nc > HCAP) {
        for_each_cell(A, S, box, [&](int c) { atomicAdd(&A.cell_count[c], 1); });
        return;
    }
    __shared__ int s_hist[HCAP];
    for (int q = threadIdx.x; q < nc; q += blockDim.x) s_hist[q] = 0;
    __syncthreads();
    for_each_cell(A, S, box, [&](int c) { atomicAdd(&s_hist[c - S.cell_off], 1); });
    __syncthreads();
    for (int q = threadIdx.x; q < nc; q += blockDim.x) {
        const int v = s_hist[q];
        if (v) atomicAdd(&A.cell_count[S.cell_off + q], v);
    }
}

// Exclusive scan of all cell counts of the batch: one block, coalesced tiles of 4096.
__global__ void __launch_bounds__(1024) cell_prefix_kernel(const int* __restrict__ cnt,
                                                           int* __restrict__ start, int n,
                                                           int* __restrict__ total)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4096) {
        const int i = base + 4 * tid;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (i + q < n) ? cnt[i + q] : 0;
        const int mine = v[0] + v[1] + v[2] + v[3];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;    // exclusive over warps
        }
        __syncthreads();
        int run = s_carry + s_warp[warp] + inc - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i + q < n) start[i + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) { start[n] = s_carry; *total = s_carry; }
}

__global__ void __launch_bounds__(256) bin_fill_kernel(BinDev A)
{
    const BinSlab S = A.slabs[slab_of_block(A.slabs, A.ns, (int)blockIdx.x)];
    const int li = ((int)blockIdx.x - S.blk_off) * (int)blockDim.x + (int)threadIdx.x;
    const int nc = S.ncx * S.ncy;
    const short4 box = (li < S.n) ? A.box[S.ob_off + li] : make_short4(1, 0, 1, 0);
    if (nc > HCAP) {
        const int ent = (A.key && li < S.n) ? A.key[S.ob_off + li] : li;
        for_each_cell(A, S, box, [&](int c) {
            const int pos = atomicAdd(&A.cell_cursor[c], 1);
            A.cell_list[A.cell_start[c] + pos] = ent;
        });
        return;
    }
    __shared__ int s_hist[HCAP];    // entries of this block per cell, then the block's cursor
    __shared__ int s_base[HCAP];    // first slot of the block in the cell's list
    for (int q = threadIdx.x; q < nc; q += blockDim.x) s_hist[q] = 0;
    __syncthreads();
    for_each_cell(A, S, box, [&](int c) { atomicAdd(&s_hist[c - S.cell_off], 1); });
    __syncthreads();
    for (int q = threadIdx.x; q < nc; q += blockDim.x) {
        const int v = s_hist[q];
        if (v) s_base[q] = A.cell_start[S.cell_off + q] + atomicAdd(&A.cell_cursor[S.cell_off + q], v);
        s_hist[q] = 0;
    }
    __syncthreads();
    const int ent = (A.key && li < S.n) ? A.key[S.ob_off + li] : li;
    for_each_cell(A, S, box, [&](int c) {
        const int q = c - S.cell_off;
        A.cell_list[s_base[q] + atomicAdd(&s_hist[q], 1)] = ent;
    });
}

// ---- per-cell sort by ob index ---------------------------------------------------------------
constexpr int NB_IN = 1024;          // value buckets of the shared-memory sorter

__device__ __forceinline__ int block_max(int v, int* s_red)
{
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = s_red[0];
    for (int w = 1; w < BIN_THREADS / 32; ++w) v = max(v, s_red[w]);
    return v;
}

// In-place exclusive prefix of s[0..n), n <= PER * BIN_THREADS; s[n] = total.  Block-wide.
template <int PER>
__device__ __forceinline__ void block_exclusive_prefix(int* s, int n, int* s_red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { const int i = tid * PER + q; v[q] = (i < n) ? s[i] : 0; sum += v[q]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) s_red[warp] = inc;
    __syncthreads();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += s_red[w];
    int total = 0;
    for (int w = 0; w < BIN_THREADS / 32; ++w) total += s_red[w];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) { const int i = tid * PER + q; if (i < n) s[i] = run; run += v[q]; }
    if (tid == 0) s[n] = total;
    __syncthreads();
}

// Sort m <= SORT_CAP distinct non-negative keys src[0..m) into s_key[0..m) (shared), block-wide.
// Ob indices reaching a cell are spread evenly over the slab, so: bucket by value (about 8
// keys per bucket, a monotone map, O(m) with shared-memory atomics), then one thread finishes
// each bucket by insertion.  A lopsided key distribution falls back to a bitonic network.
__device__ void block_sort_to_smem(const int* __restrict__ src, int m, int* __restrict__ s_key,
                                   int* __restrict__ s_bkt, int* __restrict__ s_red)
{
    const int tid = threadIdx.x;
    if (m <= 96) {       // rank by counting
        for (int t = tid; t < m; t += BIN_THREADS) {
            const int key = src[t];
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += (src[j] < key) ? 1 : 0;
            s_key[rank] = key;
        }
        __syncthreads();
        return;
    }
    int kmax = 0, kminn = 0;       // max of key and of -key
    for (int t = tid; t < m; t += BIN_THREADS) { const int k = src[t]; kmax = max(kmax, k); kminn = max(kminn, 0x7fffffff - k); }
    kmax = block_max(kmax, s_red);
    const int kmin = 0x7fffffff - block_max(kminn, s_red);
    const int nb = min(NB_IN, max(1, m / 8));
    const float scale = (float)nb / ((float)(kmax - kmin) + 1.0f);
    auto bucket = [&](int key) { return min(nb - 1, (int)((float)(key - kmin) * scale)); };
    for (int q = tid; q <= nb; q += BIN_THREADS) s_bkt[q] = 0;
    __syncthreads();
    for (int t = tid; t < m; t += BIN_THREADS) atomicAdd(&s_bkt[bucket(src[t])], 1);
    __syncthreads();
    block_exclusive_prefix<NB_IN / BIN_THREADS>(s_bkt, nb, s_red);
    for (int t = tid; t < m; t += BIN_THREADS) {
        const int key = src[t];
        s_key[atomicAdd(&s_bkt[bucket(key)], 1)] = key;
    }
    __syncthreads();
    // s_bkt[b] is now the end of bucket b
    bool lopsided = false;
    for (int bk = tid; bk < nb; bk += BIN_THREADS) {
        const int lo = bk ? s_bkt[bk - 1] : 0, hi = s_bkt[bk];
        if (hi - lo > 64) { lopsided = true; continue; }
        for (int i = lo + 1; i < hi; ++i) {
            const int key = s_key[i];
            int j = i - 1;
            while (j >= lo && s_key[j] > key) { s_key[j + 1] = s_key[j]; --j; }
            s_key[j + 1] = key;
        }
    }
    if (__syncthreads_or(lopsided ? 1 : 0)) {
        int P = 1;
        while (P < m) P <<= 1;
        for (int t = m + tid; t < P; t += BIN_THREADS) s_key[t] = 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (P >> 1); t += BIN_THREADS) {
                    const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int hi = lo | j;
                    const bool up = ((lo & k) == 0);
                    const int a = s_key[lo], b = s_key[hi];
                    if ((a > b) == up) { s_key[lo] = b; s_key[hi] = a; }
                }
                __syncthreads();
            }
        }
    }
}

// Every cell list into ascending ob index -- the reference's summation order.  Most lists are
// short (tens of obs): one warp ranks one of those by counting; the long ones are queued for
// the block-wide sorter below.
constexpr int SMALL_LIST = 256;
__global__ void __launch_bounds__(BIN_THREADS) cell_sort_small_kernel(BinDev A, int ncells)
{
    __shared__ int s_k[BIN_THREADS / 32][SMALL_LIST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * (BIN_THREADS / 32) + warp;
    if (c >= ncells) return;
    const int L = A.cell_count[c];
    if (L == 0) return;
    if (L > SMALL_LIST) {
        if (lane == 0) A.big_cells[atomicAdd(A.n_big, 1)] = c;
        return;
    }
    const int* __restrict__ in = A.cell_list + A.cell_start[c];
    int* __restrict__ out = A.cell_sorted + A.cell_start[c];
    int* k = s_k[warp];
    if (L <= 24) {          // rank by counting
        for (int t = lane; t < L; t += 32) k[t] = in[t];
        __syncwarp();
        for (int t = lane; t < L; t += 32) {
            const int key = k[t];
            int rank = 0;
            for (int j = 0; j < L; ++j) rank += (k[j] < key) ? 1 : 0;
            out[rank] = key;
        }
        return;
    }
    // bitonic network over the warp's shared-memory row, padded to a power of two
    int P = 32;
    while (P < L) P <<= 1;
    for (int t = lane; t < P; t += 32) k[t] = (t < L) ? in[t] : 0x7fffffff;
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const bool up = ((lo & kk) == 0);
                const int a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncwarp();
        }
    }
    for (int t = lane; t < L; t += 32) out[t] = k[t];
}

__global__ void __launch_bounds__(BIN_THREADS) cell_sort_kernel(BinDev A)
{
  for (int w = blockIdx.x; w < *A.n_big; w += gridDim.x) {
    __syncthreads();
    const int c = A.big_cells[w];
    const int L = A.cell_count[c];
    const int* __restrict__ in = A.cell_list + A.cell_start[c];
    int* __restrict__ out = A.cell_sorted + A.cell_start[c];
    const int tid = threadIdx.x;
    extern __shared__ int s_key[];
    __shared__ int s_bkt[NB_IN + 1];
    __shared__ int s_red[BIN_THREADS / 32];
    if (L <= SORT_CAP) {
        block_sort_to_smem(in, L, s_key, s_bkt, s_red);
        for (int t = tid; t < L; t += BIN_THREADS) out[t] = s_key[t];
        continue;
    }
    // Longer than the shared-memory sorter: bucket the keys by value in global memory first
    // (about 1024 per bucket), then sort every bucket in shared memory.
    __shared__ int s_hist[NB_MAX + 1];
    int kmax = 0;
    for (int t = tid; t < L; t += BIN_THREADS) kmax = max(kmax, in[t]);
    kmax = block_max(kmax, s_red);
    const int nb = min(NB_MAX, max(1, L / 1024));
    const float scale = (float)nb / ((float)kmax + 1.0f);
    auto bucket = [&](int key) { return min(nb - 1, (int)((float)key * scale)); };
    for (int q = tid; q <= nb; q += BIN_THREADS) s_hist[q] = 0;
    __syncthreads();
    for (int t = tid; t < L; t += BIN_THREADS) atomicAdd(&s_hist[bucket(in[t])], 1);
    __syncthreads();
    block_exclusive_prefix<NB_MAX / BIN_THREADS>(s_hist, nb, s_red);
    // scatter into the buckets (any order inside a bucket); s_key doubles as the cursors
    for (int q = tid; q < nb; q += BIN_THREADS) s_key[q] = 0;
    __syncthreads();
    for (int t = tid; t < L; t += BIN_THREADS) {
        const int key = in[t];
        const int bk = bucket(key);
        out[s_hist[bk] + atomicAdd(&s_key[bk], 1)] = key;
    }
    __syncthreads();
    int* tmp = const_cast<int*>(in);      // the fill-order list is free from here on
    for (int bk = 0; bk < nb; ++bk) {
        const int b0 = s_hist[bk], m = s_hist[bk + 1] - b0;
        if (m <= 1) continue;
        __syncthreads();
        if (m > SORT_CAP) {
            // cannot happen for evenly spread indices; kept correct: rank by counting
            for (int t = tid; t < m; t += BIN_THREADS) {
                const int key = out[b0 + t];
                int rank = 0;
                for (int j = 0; j < m; ++j) rank += (out[b0 + j] < key) ? 1 : 0;
                tmp[b0 + rank] = key;
            }
            __syncthreads();
            for (int t = tid; t < m; t += BIN_THREADS) out[b0 + t] = tmp[b0 + t];
            continue;
        }
        for (int t = tid; t < m; t += BIN_THREADS) tmp[b0 + t] = out[b0 + t];
        __syncthreads();
        block_sort_to_smem(tmp + b0, m, s_key, s_bkt, s_red);
        for (int t = tid; t < m; t += BIN_THREADS) out[b0 + t] = s_key[t];
    }
  }
}

} // namespace

int bin_build(og_ctx* ctx, const char* tag, const std::vector<BinSlab>& slabs, const short4* d_box,
              bool sorted, BinLists* out, int own_rank, int own_world, const int* key)
{
    const int ns = (int)slabs.size();
    cudaStream_t st = ctx->stream;
    const std::string t(tag);
    long long ncells = 0;
    int max_n = 0;
    for (const BinSlab& s : slabs) {
        if (s.cell_off != (int)ncells) { set_error("bin_build: cell offsets are not cumulative"); return OG_ERR_INVALID; }
        ncells += (long long)s.ncx * s.ncy;
        max_n = std::max(max_n, s.n);
    }
    if (ncells > 0x7fff0000LL) { set_error("batch too large: slabs x cells overflows"); return OG_ERR_INVALID; }
    BinDev A{};
    BinSlab* d_slabs = nullptr;
    OG_TRY(ctx->reserve_t((t + ".slabs").c_str(), (size_t)std::max(ns, 1), &d_slabs));
    OG_TRY(ctx->reserve_t((t + ".count").c_str(), (size_t)ncells + 1, &A.cell_count));
    OG_TRY(ctx->reserve_t((t + ".cursor").c_str(), (size_t)ncells + 1, &A.cell_cursor));
    OG_TRY(ctx->reserve_t((t + ".start").c_str(), (size_t)ncells + 1, &A.cell_start));
    A.slabs = d_slabs; A.ns = ns; A.box = d_box;
    A.own_rank = own_rank; A.own_world = own_world; A.key = key;
    std::vector<BinSlab> hs(slabs);
    int nblk = 0;
    for (BinSlab& q : hs) { q.blk_off = nblk; nblk += (q.n + 255) / 256; }
    OG_CUDA(cudaMemsetAsync(A.cell_count, 0, sizeof(int) * ((size_t)ncells + 1), st));
    OG_CUDA(cudaMemsetAsync(A.cell_start, 0, sizeof(int) * ((size_t)ncells + 1), st));
    int list_total = 0;
    if (ns > 0 && max_n > 0 && ncells > 0) {
        OG_TRY(ctx->put_small(d_slabs, hs.data(), sizeof(BinSlab) * ns));
        OG_CUDA(cudaMemsetAsync(A.cell_cursor, 0, sizeof(int) * ((size_t)ncells + 1), st));
        bin_count_kernel<<<nblk, 256, 0, st>>>(A);
        cell_prefix_kernel<<<1, 1024, 0, st>>>(A.cell_count, A.cell_start, (int)ncells, ctx->d_flag);
        ctx->count_launch(2);
        void* h_total = nullptr;
        OG_TRY(ctx->get_small(ctx->d_flag, sizeof(int), &h_total));
        OG_CUDA(cudaStreamSynchronize(st));
        list_total = *(const int*)h_total;
        if (list_total < 0) { set_error("cell lists overflow 2^31 entries; split the batch"); return OG_ERR_INVALID; }
    }
    const size_t nlist = (size_t)std::max(list_total, 1);
    OG_TRY(ctx->reserve_t((t + ".list").c_str(), nlist, &A.cell_list));
    A.cell_sorted = A.cell_list;
    if (sorted) OG_TRY(ctx->reserve_t((t + ".sorted").c_str(), nlist, &A.cell_sorted));
    if (list_total > 0) {
        bin_fill_kernel<<<nblk, 256, 0, st>>>(A);
        ctx->count_launch();
        if (sorted) {
            OG_TRY(ctx->reserve_t((t + ".big").c_str(), (size_t)ncells + 1, &A.big_cells));
            A.n_big = ctx->d_flag + 4;
            OG_CUDA(cudaMemsetAsync(A.n_big, 0, sizeof(int), st));
            cell_sort_small_kernel<<<(unsigned)((ncells + BIN_THREADS / 32 - 1) / (BIN_THREADS / 32)), BIN_THREADS, 0, st>>>(A, (int)ncells);
            // static (bucket tables) + dynamic (sort buffer) shared memory exceed the 48 KB default
            OG_CUDA(cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SORT_CAP * (int)sizeof(int)));
            const unsigned nblk = (unsigned)std::min<long long>(ncells, (long long)ctx->sm_count * 4);
            cell_sort_kernel<<<nblk, BIN_THREADS, SORT_CAP * sizeof(int), st>>>(A);
            ctx->count_launch(2);
        }
        OG_CUDA(cudaGetLastError());
    }
    out->slabs = d_slabs;
    out->cell_count = A.cell_count;
    out->cell_start = A.cell_start;
    out->cell_list = A.cell_sorted;
    out->total_cells = (int)ncells;
    out->list_total = list_total;
    return OG_OK;
}

} // namespace og
