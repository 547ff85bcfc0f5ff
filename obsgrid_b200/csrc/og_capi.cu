// og_capi.cu -- the extern "C" surface of libobsgrid_gpu.so (include/obsgrid_gpu.h), part 1:
// context and the per-slab drop-ins for the reference's call sites.  The per-time batched
// drivers that mirror the level x variable loops of proc_qc / proc_oa are in og_time.cu.
#include "og_ctx.h"
#include "og_device.cuh"
#include "og_capi_util.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <new>

using namespace og;

// =========================================================================================
// context
// =========================================================================================

extern "C" int og_abi_version(void) { return OG_ABI_VERSION; }
extern "C" const char* og_last_error(void) { return og::last_error(); }

extern "C" int og_init(og_ctx** out, int device)
{
    if (!out) { set_error("og_init: null ctx pointer"); return OG_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("og_init: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
        return OG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { set_error("og_init: device %d of %d", device, ndev); return OG_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        set_error("og_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return OG_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("og_init: cudaSetDevice failed"); return OG_ERR_NO_DEVICE; }
    og_ctx* c = new (std::nothrow) og_ctx();
    if (!c) return OG_ERR_ALLOC;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_out, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_a) != cudaSuccess || cudaEventCreate(&c->ev_b) != cudaSuccess ||
        cudaMalloc(&c->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&c->d_flag, 64 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&c->h_pinned_small, 4096) != cudaSuccess) {
        set_error("og_init: resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return OG_ERR_CUDA;
    }
    cudaMemset(c->d_counters, 0, 8 * sizeof(unsigned long long));
    *out = c;
    return OG_OK;
}

extern "C" int og_finalize(og_ctx* c)
{
    OG_BIND(c);
    if (!c) return OG_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->bufs) if (kv.second.p) cudaFree(kv.second.p);
    for (auto& t : c->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    cudaFree(c->d_counters); cudaFree(c->d_flag); cudaFreeHost(c->h_pinned_small);
    if (c->arena) cudaFreeHost(c->arena);
    cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b);
    for (auto& e : c->pipe_events) cudaEventDestroy(e);
    for (auto& sl : c->slots) {
        if (sl.up) cudaEventDestroy(sl.up);
        if (sl.comp) cudaEventDestroy(sl.comp);
        if (sl.down) cudaEventDestroy(sl.down);
    }
    if (c->stream_in) cudaStreamDestroy(c->stream_in);
    if (c->stream_out) cudaStreamDestroy(c->stream_out);
    if (c->stream_aux) cudaStreamDestroy(c->stream_aux);
    cudaStreamDestroy(c->stream);
    delete c;
    return OG_OK;
}

static int fetch_counters(og_ctx* c)
{
    unsigned long long h[6];
    OG_CUDA(cudaMemcpyAsync(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.cressman_updates = (int64_t)h[0];
    c->stats.cressman_candidates = (int64_t)h[1];
    c->stats.buddy_pairs_examined = (int64_t)h[3];
    c->stats.cressman_updates_ellipse = (int64_t)h[4];
    c->stats.cressman_updates_banana = (int64_t)h[5];
    return OG_OK;
}

extern "C" int og_get_stats(og_ctx* c, og_stats* out)
{
    OG_BIND(c);
    if (!c || !out) return OG_ERR_INVALID;
    OG_TRY(fetch_counters(c));
    *out = c->stats;
    return OG_OK;
}
extern "C" int og_reset_stats(og_ctx* c)
{
    OG_BIND(c);
    if (!c) return OG_ERR_INVALID;
    c->stats = og_stats{};
    OG_CUDA(cudaMemsetAsync(c->d_counters, 0, 6 * sizeof(unsigned long long), c->stream));
    return OG_OK;
}
extern "C" int og_set_counting(og_ctx* c, int on) { if (!c) return OG_ERR_INVALID; c->counting = on; return OG_OK; }
extern "C" int og_set_arithmetic(og_ctx* c, int mode)
{
    OG_BIND(c);
    if (!c || (mode != OG_ARITH_EXACT && mode != OG_ARITH_CONTRACTED)) { set_error("og_set_arithmetic: bad argument"); return OG_ERR_INVALID; }
    c->arithmetic = mode;
    return OG_OK;
}
extern "C" int og_get_arithmetic(og_ctx* c) { return c ? c->arithmetic : OG_ERR_INVALID; }
extern "C" int og_set_timing(og_ctx* c, int on) { if (!c) return OG_ERR_INVALID; c->time_kernels = on != 0; c->timers_used = 0; return OG_OK; }
extern "C" void* og_stream(og_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int og_sync(og_ctx* c) { if (!c) return OG_ERR_INVALID; OG_BIND(c); OG_CUDA(cudaStreamSynchronize(c->stream)); return OG_OK; }

extern "C" int og_last_kernel_ms(og_ctx* c, float* cressman_ms, float* buddy_ms)
{
    OG_BIND(c);
    if (!c) return OG_ERR_INVALID;
    OG_CUDA(cudaStreamSynchronize(c->stream));
    float acc[2] = {0.f, 0.f};
    for (size_t i = 0; i < c->timers_used; ++i) {
        float ms = 0.f;
        OG_CUDA(cudaEventElapsedTime(&ms, c->timers[i].a, c->timers[i].b));
        acc[c->timers[i].kind & 1] += ms;
    }
    c->timers_used = 0;
    if (cressman_ms) *cressman_ms = acc[0];
    if (buddy_ms) *buddy_ms = acc[1];
    return OG_OK;
}
extern "C" int og_fp32_peak(og_ctx* c, double* v) { if (!c || !v) return OG_ERR_INVALID; OG_BIND(c); return fp32_peak(c, v); }
extern "C" int og_selftest_div(og_ctx* c, int n, unsigned seed, int rmax, long long* bad)
{
    OG_BIND(c);
    if (!c || !bad || n <= 0 || rmax <= 0) return OG_ERR_INVALID;
    return div_selftest(c, n, seed, rmax, bad);
}
extern "C" int og_set_exchange(og_ctx* c, int rank, int world, og_exchange_fn fn, void* user)
{
    if (!c) return OG_ERR_INVALID;
    if (world <= 1 || !fn) { c->shard_rank = 0; c->shard_world = 1; c->exchange = nullptr; c->exchange_user = nullptr; return OG_OK; }
    if (rank < 0 || rank >= world) { set_error("og_set_exchange: rank %d of %d", rank, world); return OG_ERR_INVALID; }
    c->shard_rank = rank; c->shard_world = world; c->exchange = fn; c->exchange_user = user;
    return OG_OK;
}
extern "C" int og_selftest_logf(og_ctx* c, const float* x, int n, float* out)
{
    OG_BIND(c);
    if (!c || !x || !out || n <= 0) return OG_ERR_INVALID;
    return logf_selftest(c, x, n, out);
}
extern "C" int og_host_alloc(void** p, size_t bytes)
{
    if (!p) return OG_ERR_INVALID;
    if (cudaMallocHost(p, bytes ? bytes : 1) != cudaSuccess) { set_error("og_host_alloc(%zu) failed", bytes); return OG_ERR_ALLOC; }
    return OG_OK;
}
extern "C" int og_host_free(void* p) { if (p) cudaFreeHost(p); return OG_OK; }

// =========================================================================================
// per-slab drop-ins (host arrays, blocking)
// =========================================================================================
// The per-slab entry points interpolate the slab at every ob they are given (bilinear reads the
// points (INT(x), INT(y)) .. +1): an ob outside [1, iew) x [1, jns), or a NaN, would be an
// out-of-bounds device read.  The reference only ever passes obs pre-filtered by proc_ob_access
// (2 <= x <= iew-1, proc_ob_access.F90:231-234); a C ABI must not depend on that.
static int check_obs_xy(const float* x, const float* y, int n, int iew, int jns, const char* fn)
{
    for (int i = 0; i < n; ++i)
        if (!(x[i] >= 1.f && x[i] < (float)iew && y[i] >= 1.f && y[i] < (float)jns)) {
            set_error("%s: observation %d at (%g, %g) lies outside the slab [1,%d) x [1,%d)", fn, i,
                      (double)x[i], (double)y[i], iew, jns);
            return OG_ERR_INVALID;
        }
    return OG_OK;
}

extern "C" int og_innovation(og_ctx* c, const float* gridded, int iew, int jns, const float* xob,
                             const float* yob, const float* obs_value, int n, float scale,
                             int use_first_guess, float* diff, float* fg_value)
{
    OG_BIND(c);
    REQUIRE(c && gridded && diff && n >= 0 && iew > 2 && jns > 2, "bad argument");
    if (n == 0) return OG_OK;
    REQUIRE(xob && yob && obs_value, "null observation array");
    OG_TRY(check_obs_xy(xob, yob, n, iew, jns, __func__));
    float *d_g, *d_x, *d_y, *d_v, *d_d, *d_f;
    OG_TRY(upload(c, "io.g", gridded, (size_t)iew * jns, &d_g));
    OG_TRY(upload(c, "io.x", xob, (size_t)n, &d_x));
    OG_TRY(upload(c, "io.y", yob, (size_t)n, &d_y));
    OG_TRY(upload(c, "io.v", obs_value, (size_t)n, &d_v));
    OG_TRY(c->reserve_t("io.diff", (size_t)n, &d_d));
    OG_TRY(c->reserve_t("io.fg", (size_t)n, &d_f));
    OG_TRY(innovation_device(c, d_g, iew, jns, d_x, d_y, d_v, n, scale, use_first_guess, d_d,
                             fg_value ? d_f : nullptr));
    OG_TRY(download(c, diff, d_d, (size_t)n));
    if (fg_value && use_first_guess) OG_TRY(download(c, fg_value, d_f, (size_t)n));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

static bool aniso(int var) { return var == OG_VAR_UU || var == OG_VAR_VV || var == OG_VAR_RH; }

extern "C" int og_cressman(og_ctx* c, const float* diff, const float* xob, const float* yob, int n,
                           float* gridded, int iew, int jns, int crsdot, int var,
                           int radius_influence, float dxd_km, const float* u_banana,
                           const float* v_banana, float pressure_hpa, int passes, int smooth_type,
                           int use_first_guess, const float* fg_value,
                           int scale_cressman_rh_decreases)
{
    OG_BIND(c);
    REQUIRE(c && gridded && n >= 0 && iew > 2 && jns > 2, "bad argument");
    REQUIRE(var >= OG_VAR_UU && var <= OG_VAR_PSFC, "unknown variable code");
    REQUIRE(radius_influence >= 0 && radius_influence < 8000, "radius out of range");
    REQUIRE(crsdot == 0 || crsdot == 1, "crsdot must be 0 or 1");
    REQUIRE(n == 0 || (diff && xob && yob), "null observation array");
    REQUIRE(!aniso(var) || (u_banana && v_banana), "UU/VV/RH need u_banana and v_banana");
    OG_TRY(check_obs_xy(xob, yob, n, iew, jns, __func__));
    const size_t npts = (size_t)iew * jns;
    float *d_g, *d_ub = nullptr, *d_vb = nullptr, *d_x, *d_y, *d_d, *d_f;
    OG_TRY(upload(c, "io.g", gridded, npts, &d_g));
    if (aniso(var)) {
        OG_TRY(upload(c, "io.ub", u_banana, npts, &d_ub));
        OG_TRY(upload(c, "io.vb", v_banana, npts, &d_vb));
    }
    OG_TRY(upload(c, "io.x", xob, (size_t)n, &d_x));
    OG_TRY(upload(c, "io.y", yob, (size_t)n, &d_y));
    OG_TRY(upload(c, "io.diff", diff, (size_t)n, &d_d));
    if (fg_value) OG_TRY(upload(c, "io.fg", fg_value, (size_t)n, &d_f));
    else { OG_TRY(c->reserve_t("io.fg", (size_t)std::max(n, 1), &d_f)); OG_CUDA(cudaMemsetAsync(d_f, 0, sizeof(float) * (size_t)std::max(n, 1), c->stream)); }

    OaBatch b;
    b.iew = iew; b.jns = jns; b.crsdot = crsdot; b.dxd = dxd_km;
    b.use_first_guess = use_first_guess; b.scale_rh = scale_cressman_rh_decreases;
    b.smooth_type = smooth_type;
    OaSlabHost s;
    s.g = d_g; s.ub = d_ub; s.vb = d_vb; s.n = n; s.ob_off = 0; s.var = var;
    s.pressure = pressure_hpa; s.scale = 1.f; s.passes = passes; s.nscan = 1;
    s.radius[0] = radius_influence;
    b.slabs.push_back(s);
    b.total_obs = n; b.xob = d_x; b.yob = d_y; b.val = nullptr; b.diff = d_d; b.fg = d_f;
    b.given_diff = true; b.clean_rh = false;
    OG_TRY(oa_run_batch(c, b));
    OG_TRY(download(c, gridded, d_g, npts));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_oa_slab(og_ctx* c, float* gridded, int iew, int jns, int var, float pressure_hpa,
                          const float* obs_value, const float* xob, const float* yob, int n,
                          float scale, const int* radius_influence,
                          float radius_influence_sfc_mult, float dxd_km, const float* u_banana,
                          const float* v_banana, int passes, int smooth_type,
                          int use_first_guess, int scale_cressman_rh_decreases, int* n_scans_done)
{
    OG_BIND(c);
    REQUIRE(c && gridded && radius_influence && n >= 0 && iew > 2 && jns > 2, "bad argument");
    REQUIRE(var >= OG_VAR_UU && var <= OG_VAR_PSFC, "unknown variable code");
    REQUIRE(n == 0 || (obs_value && xob && yob), "null observation array");
    REQUIRE(!aniso(var) || (u_banana && v_banana), "UU/VV/RH need u_banana and v_banana");
    OG_TRY(check_obs_xy(xob, yob, n, iew, jns, __func__));
    const size_t npts = (size_t)iew * jns;
    OaBatch b;
    OaSlabHost s;
    int rrc = scan_radii(radius_influence, radius_influence_sfc_mult, pressure_hpa, s.radius, &s.nscan);
    float *d_g, *d_ub = nullptr, *d_vb = nullptr, *d_x, *d_y, *d_v;
    OG_TRY(upload(c, "io.g", gridded, npts, &d_g));
    if (aniso(var)) {
        OG_TRY(upload(c, "io.ub", u_banana, npts, &d_ub));
        OG_TRY(upload(c, "io.vb", v_banana, npts, &d_vb));
    }
    OG_TRY(upload(c, "io.x", xob, (size_t)n, &d_x));
    OG_TRY(upload(c, "io.y", yob, (size_t)n, &d_y));
    OG_TRY(upload(c, "io.v", obs_value, (size_t)n, &d_v));
    b.iew = iew; b.jns = jns; b.crsdot = 1; b.dxd = dxd_km;
    b.use_first_guess = use_first_guess; b.scale_rh = scale_cressman_rh_decreases;
    b.smooth_type = smooth_type;
    s.g = d_g; s.ub = d_ub; s.vb = d_vb; s.n = n; s.ob_off = 0; s.var = var;
    s.pressure = pressure_hpa; s.scale = scale; s.passes = passes;
    b.slabs.push_back(s);
    b.total_obs = n; b.xob = d_x; b.yob = d_y; b.val = d_v;
    b.given_diff = false; b.clean_rh = true;
    OG_TRY(oa_run_batch(c, b));
    OG_TRY(download(c, gridded, d_g, npts));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    if (n_scans_done) *n_scans_done = s.nscan;
    if (rrc != OG_OK) set_error("surface scan-1 radius below 4.5 cells (proc_oa.F90:694-703)");
    return rrc;
}

extern "C" int og_local_time(og_ctx* c, int time_hhmm, const float* lonob, int n, int* local_time)
{
    OG_BIND(c);
    REQUIRE(c && n >= 0 && (n == 0 || (lonob && local_time)), "bad argument");
    if (n == 0) return OG_OK;
    float* d_l; int* d_o;
    OG_TRY(upload(c, "io.lon", lonob, (size_t)n, &d_l));
    OG_TRY(c->reserve_t("io.lt", (size_t)n, &d_o));
    OG_TRY(local_time_device(c, time_hhmm, d_l, n, d_o));
    OG_TRY(download(c, local_time, d_o, (size_t)n));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_fg_t_at_point(og_ctx* c, const float* fg_3d_t, const float* fg_3d_h, int jns,
                                int iew, int kbu, const float* x_location, const float* y_location,
                                const float* height_msl, int n, float* fg_t)
{
    OG_BIND(c);
    REQUIRE(c && fg_3d_t && fg_3d_h && n >= 0 && iew > 2 && jns > 2 && kbu >= 3, "bad argument");
    if (n == 0) return OG_OK;
    REQUIRE(x_location && y_location && height_msl && fg_t, "null point array");
    const size_t vol = (size_t)iew * jns * kbu;
    float *d_t, *d_h, *d_x, *d_y, *d_z, *d_o;
    OG_TRY(upload(c, "io.t3", fg_3d_t, vol, &d_t));
    OG_TRY(upload(c, "io.h3", fg_3d_h, vol, &d_h));
    OG_TRY(upload(c, "io.x", x_location, (size_t)n, &d_x));
    OG_TRY(upload(c, "io.y", y_location, (size_t)n, &d_y));
    OG_TRY(upload(c, "io.z", height_msl, (size_t)n, &d_z));
    OG_TRY(c->reserve_t("io.fgt", (size_t)n, &d_o));
    OG_TRY(fg_t_at_point_device(c, d_t, d_h, jns, iew, kbu, d_x, d_y, d_z, n, d_o));
    OG_TRY(download(c, fg_t, d_o, (size_t)n));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_ob_density(og_ctx* c, const float* xob, const float* yob, float grid_dist, int n,
                             float* tobbox, int iew, int jns)
{
    OG_BIND(c);
    REQUIRE(c && tobbox && n >= 0 && iew > 2 && jns > 2 && iew <= 32000 && jns <= 32000, "bad argument");
    REQUIRE(grid_dist > 0.f, "grid_dist must be positive");
    if (n == 0) return OG_OK;
    REQUIRE(xob && yob, "null observation array");
    const size_t npts = (size_t)iew * jns;
    float *d_x, *d_y, *d_t;
    OG_TRY(upload(c, "io.x", xob, (size_t)n, &d_x));
    OG_TRY(upload(c, "io.y", yob, (size_t)n, &d_y));
    OG_TRY(upload(c, "io.tobbox", tobbox, npts, &d_t));
    OG_TRY(ob_density_device(c, d_x, d_y, grid_dist, n, d_t, iew, jns));
    OG_TRY(download(c, tobbox, d_t, npts));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_obs_distance(og_ctx* c, const float* tobbox, float* distance, int iew, int jns, float dx)
{
    OG_BIND(c);
    REQUIRE(c && tobbox && distance && iew > 0 && jns > 0 && iew <= 32000 && jns <= 32000, "bad argument");
    const size_t npts = (size_t)iew * jns;
    float *d_t, *d_d;
    OG_TRY(upload(c, "io.tobbox", tobbox, npts, &d_t));
    OG_TRY(c->reserve_t("io.odis", npts, &d_d));
    OG_TRY(obs_distance_device(c, d_t, d_d, iew, jns, dx));
    OG_TRY(download(c, distance, d_d, npts));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

// ---- post-analysis epilogue -------------------------------------------
extern "C" int og_wind_increment_to_c_grid(og_ctx* c, float* u, const float* uA, const float* uC,
                                           float* v, const float* vA, const float* vC, int jns,
                                           int iew, int kbu)
{
    OG_BIND(c);
    REQUIRE(c && jns >= 2 && iew >= 2 && kbu >= 1 && kbu <= 65535 && iew <= 65535, "bad argument");
    REQUIRE((u && uA && uC) || (v && vA && vC), "null field");
    const size_t n = (size_t)jns * iew * kbu;
    for (int dir = 0; dir < 2; ++dir) {
        float* w = dir ? v : u;
        const float* wA = dir ? vA : uA;
        const float* wC = dir ? vC : uC;
        if (!w) continue;
        REQUIRE(wA && wC, "null A-grid or C-grid wind");
        float *d_w, *d_a, *d_c, *d_o;
        OG_TRY(upload(c, "fin.w", w, n, &d_w));
        OG_TRY(upload(c, "fin.wA", wA, n, &d_a));
        OG_TRY(upload(c, "fin.wC", wC, n, &d_c));
        OG_TRY(c->reserve_t("fin.out", n, &d_o));
        OG_TRY(wind_increment_device(c, dir, d_w, d_a, d_c, d_o, jns, iew, kbu));
        OG_TRY(download(c, w, d_o, n));
        OG_CUDA(cudaStreamSynchronize(c->stream));
    }
    return OG_OK;
}

extern "C" int og_temporal_interp(og_ctx* c, const float* first, const float* second, float* out,
                                  int jns, int iew, int nlev, int icount_fdda, int icount_1,
                                  int icount_2, int cross)
{
    OG_BIND(c);
    REQUIRE(c && first && second && out && jns >= 2 && iew >= 2 && nlev >= 1, "bad argument");
    REQUIRE(iew <= 65535 && nlev <= 65535, "too many columns or levels");
    // tinterp_module.F90:98-121: integer weights, the both-zero special case, the divide check
    int w1 = icount_2 - icount_fdda, w2 = icount_fdda - icount_1;
    if (w1 == 0 && w2 == 0) w1 = 1;
    if (w1 + w2 == 0) { set_error("temporal_interp: weights sum to zero (tinterp_module.F90:112)"); return OG_ERR_INVALID; }
    const size_t n = (size_t)jns * iew * nlev;
    float *d_a, *d_b, *d_o;
    OG_TRY(upload(c, "fin.w", first, n, &d_a));
    OG_TRY(upload(c, "fin.wA", second, n, &d_b));
    OG_TRY(c->reserve_t("fin.out", n, &d_o));
    OG_TRY(temporal_interp_device(c, d_a, d_b, d_o, jns, iew, nlev, (float)w1, (float)w2, (float)(w1 + w2), cross));
    OG_TRY(download(c, out, d_o, n));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

// device-resident form: first / second / out are device pointers (the surface-FDDA chain keeps the
// analysed 3-D times in HBM between its phases)
extern "C" int og_dev_temporal_interp(og_ctx* c, const float* first, const float* second, float* out,
                                      int jns, int iew, int nlev, int icount_fdda, int icount_1,
                                      int icount_2, int cross)
{
    OG_BIND(c);
    REQUIRE(c && first && second && out && jns >= 2 && iew >= 2 && nlev >= 1, "bad argument");
    REQUIRE(iew <= 65535 && nlev <= 65535, "too many columns or levels");
    int w1 = icount_2 - icount_fdda, w2 = icount_fdda - icount_1;
    if (w1 == 0 && w2 == 0) w1 = 1;
    if (w1 + w2 == 0) { set_error("temporal_interp: weights sum to zero (tinterp_module.F90:112)"); return OG_ERR_INVALID; }
    return temporal_interp_device(c, first, second, out, jns, iew, nlev, (float)w1, (float)w2, (float)(w1 + w2), cross);
}

extern "C" int og_pres_sfc_from_slp(og_ctx* c, const float* pres_sfc, const float* slp_C,
                                    const float* slp_x, float* out, int jns, int iew)
{
    OG_BIND(c);
    REQUIRE(c && pres_sfc && slp_C && slp_x && out && jns > 0 && iew > 0, "bad argument");
    const size_t n = (size_t)jns * iew;
    float *d_p, *d_c, *d_x, *d_o;
    OG_TRY(upload(c, "fin.w", pres_sfc, n, &d_p));
    OG_TRY(upload(c, "fin.wA", slp_C, n, &d_c));
    OG_TRY(upload(c, "fin.wC", slp_x, n, &d_x));
    OG_TRY(c->reserve_t("fin.out", n, &d_o));
    OG_TRY(pres_sfc_device(c, d_p, d_c, d_x, d_o, n));
    OG_TRY(download(c, out, d_o, n));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_mixing_ratio(og_ctx* c, float* rh, const float* t, const float* h,
                               const float* terrain, const float* slp_x, const float* pressure,
                               int iew, int jns, int kbu, float* qv, float* psfc)
{
    OG_BIND(c);
    REQUIRE(c && rh && t && h && terrain && slp_x && pressure && qv && psfc, "null argument");
    REQUIRE(iew >= 2 && jns >= 2 && kbu >= 2 && iew <= 65535, "bad dimensions");
    const size_t per = (size_t)jns * iew, n = per * kbu;
    float *d_rh, *d_t, *d_h, *d_ter, *d_slp, *d_p, *d_qv, *d_ps;
    OG_TRY(upload(c, "fin.w", rh, n, &d_rh));
    OG_TRY(upload(c, "fin.wA", t, n, &d_t));
    OG_TRY(upload(c, "fin.wC", h, n, &d_h));
    OG_TRY(upload(c, "fin.ter", terrain, per, &d_ter));
    OG_TRY(upload(c, "fin.slp", slp_x, per, &d_slp));
    OG_TRY(upload(c, "fin.p", pressure, (size_t)kbu, &d_p));
    // qv / psfc are only written on the cross points (1:jns-1, 1:iew-1): keep the caller's values elsewhere
    OG_TRY(upload(c, "fin.out", qv, n, &d_qv));
    OG_TRY(upload(c, "fin.ps", psfc, per, &d_ps));
    OG_TRY(mixing_ratio_device(c, d_rh, d_t, d_h, d_ter, d_slp, d_p, pressure, iew, jns, kbu, d_qv, d_ps));
    OG_TRY(download(c, rh, d_rh, n));
    OG_TRY(download(c, qv, d_qv, n));
    OG_TRY(download(c, psfc, d_ps, per));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

// shared body of og_error_max / og_buddy_check / og_qc_slab
static int qc_slab_common(og_ctx* c, int do_emax, int do_buddy, const float* gridded, int iew,
                          int jns, int var, float pressure, const float* obs, const float* xob,
                          const float* yob, const float* elev, const int* local_time, int n,
                          float max_error, float max_buddy, float buddy_weight, float dx_km,
                          int* qc_flag, float* aob, float* err_max, float* thr_max,
                          float* err_buddy, float* difmax, int* bnf)
{
    REQUIRE(c && gridded && n >= 0 && iew > 2 && jns > 2, "bad argument");
    REQUIRE(var >= OG_VAR_UU && var <= OG_VAR_DEWPOINT, "unknown variable code");
    if (n == 0) return OG_OK;
    REQUIRE(obs && xob && yob && qc_flag, "null observation array");
    REQUIRE(var != OG_VAR_PMSLPSFC || elev, "PMSLPSFC needs station_elevation");
    REQUIRE(!(do_emax && var == OG_VAR_TT && pressure > 1000.1f) || local_time, "surface TT needs local_time");
    OG_TRY(check_obs_xy(xob, yob, n, iew, jns, "og_qc_slab / og_error_max / og_buddy_check"));
    const size_t nn = (size_t)n;
    float *d_g, *d_o, *d_x, *d_y, *d_e = nullptr; int *d_lt = nullptr, *d_q;
    OG_TRY(upload(c, "io.g", gridded, (size_t)iew * jns, &d_g));
    OG_TRY(upload(c, "io.v", obs, nn, &d_o));
    OG_TRY(upload(c, "io.x", xob, nn, &d_x));
    OG_TRY(upload(c, "io.y", yob, nn, &d_y));
    if (elev) OG_TRY(upload(c, "io.elev", elev, nn, &d_e));
    else { OG_TRY(c->reserve_t("io.elev", nn, &d_e)); OG_CUDA(cudaMemsetAsync(d_e, 0, nn * 4, c->stream)); }
    if (local_time) OG_TRY(upload(c, "io.lt", local_time, nn, &d_lt));
    else { OG_TRY(c->reserve_t("io.lt", nn, &d_lt)); OG_CUDA(cudaMemsetAsync(d_lt, 0, nn * 4, c->stream)); }
    OG_TRY(upload(c, "io.qc", qc_flag, nn, &d_q));
    QcBatch b;
    b.iew = iew; b.jns = jns; b.dx_km = dx_km; b.buddy_weight = buddy_weight;
    b.do_error_max = do_emax; b.do_buddy = do_buddy;
    QcSlabHost s;
    s.g = d_g; s.n = n; s.ob_off = 0; s.var = var; s.pressure = pressure;
    s.max_error = max_error; s.max_buddy = max_buddy;
    b.slabs.push_back(s);
    b.total_obs = n; b.xob = d_x; b.yob = d_y; b.val = d_o; b.elev = d_e; b.local_time = d_lt;
    b.qc = d_q;
    if (aob) OG_TRY(c->reserve_t("io.d_aob", nn, &b.aob));
    if (err_max) OG_TRY(c->reserve_t("io.d_em", nn, &b.err_max));
    if (thr_max) OG_TRY(c->reserve_t("io.d_tm", nn, &b.thr_max));
    if (err_buddy) OG_TRY(c->reserve_t("io.d_eb", nn, &b.err_buddy));
    if (difmax) OG_TRY(c->reserve_t("io.d_dm", nn, &b.difmax));
    if (bnf) OG_TRY(c->reserve_t("io.d_bnf", nn, &b.buddy_num_final));
    OG_TRY(qc_run_batch(c, b));
    OG_TRY(download(c, qc_flag, d_q, nn));
    OG_TRY(download(c, aob, b.aob, nn));
    OG_TRY(download(c, err_max, b.err_max, nn));
    OG_TRY(download(c, thr_max, b.thr_max, nn));
    OG_TRY(download(c, err_buddy, b.err_buddy, nn));
    OG_TRY(download(c, difmax, b.difmax, nn));
    OG_TRY(download(c, bnf, b.buddy_num_final, nn));
    OG_CUDA(cudaStreamSynchronize(c->stream));
    return OG_OK;
}

extern "C" int og_error_max(og_ctx* c, const float* obs, const float* xob, const float* yob,
                            const float* station_elevation, int* qc_flag, int n,
                            const float* gridded, int iew, int jns, int var, float max_difference,
                            float pressure_hpa, const int* local_time, float* aob, float* err,
                            float* thr)
{
    OG_BIND(c);
    return qc_slab_common(c, 1, 0, gridded, iew, jns, var, pressure_hpa, obs, xob, yob,
                          station_elevation, local_time, n, max_difference, 0.f, 1.f, 1.f, qc_flag,
                          aob, err, thr, nullptr, nullptr, nullptr);
}

extern "C" int og_buddy_check(og_ctx* c, const float* obs, int n, const float* xob, const float* yob,
                              int* qc_flag, const float* station_elevation, const float* gridded,
                              int iew, int jns, int var, float pressure_hpa, float dx_km,
                              float buddy_weight, float buddy_difference, float* aob, float* err,
                              float* difmax, int* buddy_num_final)
{
    OG_BIND(c);
    return qc_slab_common(c, 0, 1, gridded, iew, jns, var, pressure_hpa, obs, xob, yob,
                          station_elevation, nullptr, n, 0.f, buddy_difference, buddy_weight, dx_km,
                          qc_flag, aob, nullptr, nullptr, err, difmax, buddy_num_final);
}

extern "C" int og_qc_slab(og_ctx* c, int do_error_max, int do_buddy, const float* gridded, int iew,
                          int jns, int var, float pressure_hpa, const float* obs, const float* xob,
                          const float* yob, const float* station_elevation, const int* local_time,
                          int n, float max_error, float max_buddy, float buddy_weight, float dx_km,
                          int* qc_flag)
{
    OG_BIND(c);
    return qc_slab_common(c, do_error_max, do_buddy, gridded, iew, jns, var, pressure_hpa, obs, xob,
                          yob, station_elevation, local_time, n, max_error, max_buddy, buddy_weight,
                          dx_km, qc_flag, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

