// og_ctx.h -- host-side context of libobsgrid_gpu.so: device, stream, grow-only buffers,
// counters and error reporting.  Internal; the public surface is include/obsgrid_gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/obsgrid_gpu.h"

namespace og {

void set_error(const char* fmt, ...);

#define OG_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            og::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                          cudaGetErrorString(e__));                                       \
            return OG_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define OG_TRY(call)                                                                      \
    do {                                                                                  \
        int rc__ = (call);                                                                \
        if (rc__ != OG_OK) return rc__;                                                   \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

} // namespace og

struct og_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;   // copy streams of og_qc_oa_time
    cudaStream_t stream_aux = nullptr;                        // second compute stream (scan kernels)
    std::vector<cudaEvent_t> pipe_events;
    struct Slot {
        cudaEvent_t up = nullptr, comp = nullptr, down = nullptr;
        bool down_pending = false;
        bool has_td = false, has_bogus = false;      // optional rows the last upload brought (ABI 3)
    };
    Slot slots[OG_SLOTS];
    og_stats stats{};
    // a slab over several GPUs (og_set_exchange): the buddy check's cells are dealt cell % world
    int shard_rank = 0, shard_world = 1;
    og_exchange_fn exchange = nullptr;
    void* exchange_user = nullptr;
    int counting = 0;
    int arithmetic = OG_ARITH_EXACT;            // og_set_arithmetic
    unsigned long long* d_counters = nullptr;   // [0] updates, [1] candidates
    int* d_flag = nullptr;                      // small device scratch (changed flags)
    int* h_pinned_small = nullptr;              // pinned host scratch (>= 4 KiB)
    std::map<std::string, og::DevBuf> bufs;     // named grow-only device buffers
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // kernel timing: event pairs recorded around the Cressman scan / buddy kernels when
    // time_kernels is on; resolved lazily by og_last_kernel_ms (no sync on the hot path)
    struct Timer { cudaEvent_t a, b; int kind; };
    std::vector<Timer> timers;
    size_t timers_used = 0;
    bool time_kernels = false;
    int timer_begin(int kind);   // records the start event; returns slot or -1
    void timer_end(int slot);

    // grow-only named device buffer; contents are NOT preserved on growth
    int reserve(const char* name, size_t bytes, void** out);
    template <class T> int reserve_t(const char* name, size_t count, T** out)
    {
        void* p = nullptr;
        int rc = reserve(name, count * sizeof(T), &p);
        *out = static_cast<T*>(p);
        return rc;
    }
    void count_launch(int n = 1) { stats.kernel_launches += n; }

    // Small host<->device transfers on the compute stream go through a mapped pinned arena and
    // a copy kernel, NOT the DMA engines: a 4-byte readback queued behind another time period's
    // 180 MB download on the (in-order) D2H engine would stall the compute stream for
    // milliseconds (og_time_* pipeline).  put_small: the source may be reused on return.
    // get_small: *host points into the arena, valid after the next stream synchronisation.
    unsigned char* arena = nullptr;
    size_t arena_size = 0, arena_pos = 0;
    int arena_take(size_t bytes, unsigned char** p);
    int put_small(void* dev_dst, const void* host_src, size_t bytes);
    int get_small(const void* dev_src, size_t bytes, void** host);
};

namespace og {

// ---- OA (og_oa.cu) -------------------------------------------------------------------
struct OaSlabHost {             // one (level, variable) slab of a batch, host description
    float* g = nullptr;         // device: field to analyse, (iew,jns) x fastest, in/out
    const float* ub = nullptr;  // device: pre-analysis u of the level (banana scheme)
    const float* vb = nullptr;
    int n = 0;                  // obs in the slab (known on host)
    int ob_off = 0;             // first ob of the slab in the batch-wide ob arrays
    int var = 0;
    float pressure = 0.f;
    float scale = 1.f;
    int passes = 0;
    int nscan = 0;
    int radius[OG_MAX_SCAN] = {0};
};

struct OaBatch {
    int iew = 0, jns = 0, crsdot = 1;
    float dxd = 0.f;
    int use_first_guess = 1, scale_rh = 0, smooth_type = 1;
    std::vector<OaSlabHost> slabs;
    int total_obs = 0;
    // device, batch-wide, length total_obs
    const float* xob = nullptr;
    const float* yob = nullptr;
    const float* val = nullptr;     // obs_value (innovations are derived), or
    float* diff = nullptr;          // innovations (given when given_diff, else scratch)
    float* fg = nullptr;            // fg_value (given when given_diff, else scratch)
    bool given_diff = false;        // per-slab og_cressman: diff / fg supplied, one scan
    bool clean_rh = true;           // apply the final RH clamp (proc_oa.F90:803)
};

int oa_run_batch(og_ctx* ctx, OaBatch& b);
int innovation_device(og_ctx* ctx, const float* g, int iew, int jns, const float* x,
                      const float* y, const float* val, int n, float scale, int use_fg,
                      float* diff, float* fg);

// ---- QC (og_qc.cu) -------------------------------------------------------------------
struct QcSlabHost {
    const float* g = nullptr;   // device: first-guess slab (iew,jns) as error_max sees it
    int n = 0;
    int ob_off = 0;
    int var = 0;
    float pressure = 0.f;
    float max_error = 0.f;      // error_difference  (proc_qc.F90:264-309)
    float max_buddy = 0.f;      // buddy_difference
};

struct QcBatch {
    int iew = 0, jns = 0;
    float dx_km = 0.f, buddy_weight = 1.f;
    int do_error_max = 1, do_buddy = 1;
    std::vector<QcSlabHost> slabs;
    int total_obs = 0;
    const float* xob = nullptr;
    const float* yob = nullptr;
    const float* val = nullptr;
    const float* elev = nullptr;
    const int* local_time = nullptr;
    int* qc = nullptr;                    // in/out flags
    // optional diagnostics (device, nullable)
    float* aob = nullptr;
    float* err_max = nullptr;
    float* thr_max = nullptr;
    float* err_buddy = nullptr;
    float* difmax = nullptr;
    int* buddy_num_final = nullptr;
};

int qc_run_batch(og_ctx* ctx, QcBatch& b);
int local_time_device(og_ctx* ctx, int time_hhmm, const float* lon, int n, int* out);
int fg_t_at_point_device(og_ctx* ctx, const float* t3, const float* h3, int jns, int iew, int kbu,
                         const float* x, const float* y, const float* h, int n, float* out);

// ---- stable binning (og_bin.cu) --------------------------------------------------------
struct BinSlab {
    int n, ob_off;      // obs of the slab in the batch-wide per-ob arrays
    int cell;           // cell edge in grid points
    int ncx, ncy;       // cells per row / column
    int cell_off;       // first cell of the slab in the per-cell arrays (cumulative)
    int blk_off;        // first 256-ob block of the slab in the flat per-ob grids (filled by bin_build)
};
struct BinLists {
    const BinSlab* slabs = nullptr;   // device copy of the slab table
    int* cell_count = nullptr;        // [total_cells+1] obs whose box touches the cell
    int* cell_start = nullptr;        // [total_cells+1] exclusive prefix over the whole batch
    int* cell_list = nullptr;         // [list_total] local ob indices; ascending when sorted
    int total_cells = 0;
    int list_total = 0;
};
// box: device, one 1-based inclusive {i0,i1,j0,j1} per ob (i0>i1 = not binned).  Synchronises
// the stream once (the list size comes back to the host to size the lists).
// own_rank / own_world: build the lists of the cells with (cell index within the slab) % own_world ==
// own_rank only (the other cells get empty lists); 0 / 1: every cell.
// key (nullable): per ob (batch-wide index) the value stored in the lists instead of the local index
// -- with sorted lists, the order of the entries (the buddy check orders a cell's targets in space
// this way); entries must stay distinct non-negative ints.
int bin_build(og_ctx* ctx, const char* tag, const std::vector<BinSlab>& slabs, const short4* d_box,
              bool sorted, BinLists* out, int own_rank = 0, int own_world = 1, const int* key = nullptr);

// ---- density / distance (og_density.cu) -------------------------------------------------
int ob_density_device(og_ctx* ctx, const float* d_x, const float* d_y, float grid_dist, int n,
                      float* d_tobbox, int iew, int jns);
int obs_distance_device(og_ctx* ctx, const float* d_tobbox, float* d_distance, int iew, int jns, float dx);

// ---- post-analysis epilogue (og_final.cu) ------------------------------------------------
int wind_increment_device(og_ctx* ctx, int dir, const float* w, const float* wA, const float* wC,
                          float* out, int jns, int iew, int kbu);
int temporal_interp_device(og_ctx* ctx, const float* a, const float* b, float* out, int jns, int iew,
                           int nlev, float w1, float w2, float wd, int cross);
int pres_sfc_device(og_ctx* ctx, const float* pres, const float* slp_C, const float* slp_x,
                    float* out, size_t n);
int mixing_ratio_device(og_ctx* ctx, float* rh, const float* t, const float* h, const float* terrain,
                        const float* slp_x, const float* d_pressure, const float* h_pressure, int iew,
                        int jns, int kbu, float* qv, float* psfc);

// ---- misc (og_misc.cu) ---------------------------------------------------------------
const char* last_error();
int div_selftest(og_ctx* ctx, int n, unsigned seed, int rmax, long long* mismatches);
int logf_selftest(og_ctx* ctx, const float* x, int n, float* out);
int fp32_peak(og_ctx* ctx, double* lane_ops_per_s);
// (jns,iew) y-fastest <-> (iew,jns) x-fastest for nlev stacked levels
int transpose_levels(og_ctx* ctx, const float* src, float* dst, int rows_fast, int cols_slow,
                     int nlev, float* dst2 = nullptr);   // dst2: optional second copy of the result

} // namespace og
