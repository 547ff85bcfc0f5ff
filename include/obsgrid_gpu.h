/*
 * obsgrid_gpu.h -- C ABI of libobsgrid_gpu.so, the B200 (sm_100a) implementation of
 * OBSGRID's objective-analysis hot path.
 *
 * The Fortran driver of wrf-model/OBSGRID keeps everything above this boundary
 * (namelist, little_r ingest, linked-list obs store, netCDF I/O).  The entry points
 * below replace, one for one, the flat-array procedures the reference's stage
 * drivers call (all file:line citations are relative to the reference tree):
 *
 *   og_innovation  <- inline loops station_loop_1fg / station_loop_5fg,
 *                     src/proc_oa.F90:517-554 and :750-778
 *   og_cressman    <- SUBROUTINE cressman, src/oa_module.F90:25-218
 *                     (dist_uu_vv_rh :229-574, dist_not_uu_vv_rh :577-627,
 *                      smooth_5 :928-1006, smoother_desmoother :875-923)
 *   og_oa_slab     <- the multi_scan loop of one (level, variable) slab,
 *                     src/proc_oa.F90:491-825 (radius scaling :663-707, clean_rh :803)
 *   og_error_max   <- SUBROUTINE error_max,   src/qc1_module.F90:9-300
 *   og_buddy_check <- SUBROUTINE buddy_check, src/qc2_module.F90:9-388
 *   og_qc_slab     <- the error_max + buddy_check pair of one slab, fused,
 *                     src/proc_qc.F90:371-392
 *   og_local_time  <- SUBROUTINE local, src/qc0_module.F90:236-309
 *   og_qc_time / og_oa_time <- the level x variable loops of proc_qc
 *                     (src/proc_qc.F90:243-408 + qc_consistency, qc0_module.F90:313-434)
 *                     and proc_oa (src/proc_oa.F90:356-849) over one analysis time,
 *                     fed from a flat SoA observation table instead of per-slab
 *                     proc_ob_access calls (src/proc_ob_access.F90:146-403).
 *
 * Conventions
 *   - plain C: pointers, ints, floats.  No torch / C++ types cross this boundary.
 *   - every array argument is HOST memory owned by the caller; the library stages it
 *     through its own pinned buffers and never keeps a pointer after the call returns.
 *     (The og_dev_* entry points at the bottom take DEVICE pointers; they exist for
 *     benchmarks and for callers that already keep the data resident.)
 *   - 2-D slabs are (iew,jns), x fastest, exactly as Fortran passes `gridded(iew,jns)`:
 *     gridded(i,j) == g[(j-1)*iew + (i-1)].  3-D first-guess fields of the per-time
 *     API are (jns,iew,kbu), y fastest (src/first_guess.inc).
 *   - xob/yob are the reference's 1-based REAL(4) dot-grid coordinates.
 *   - Fortran LOGICALs are passed as int 0/1; CHARACTER(8) variable names as og_var codes.
 *   - all functions return OG_OK (0) or a negative og_status; they never abort/STOP.
 *     og_last_error() returns a message for the last failure on the calling thread.
 *   - there is NO CPU fallback: og_init fails with OG_ERR_NO_DEVICE without an
 *     sm_100 GPU.
 *   - callers are single threaded per og_ctx (like the reference); calls block.
 *     Contexts are independent: several may live on one GPU or on several, each driven by
 *     its own host thread (two contexts on one GPU overlap the list-size round trips of one
 *     time period with the kernels of another).  Every entry point makes its context's
 *     device current, so a context may be used from any host thread.
 */
#ifndef OBSGRID_GPU_H
#define OBSGRID_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OG_ABI_VERSION 3

typedef enum og_status {
    OG_OK = 0,
    OG_ERR_NO_DEVICE = -1,   /* no sm_100 device / CUDA runtime failure at init        */
    OG_ERR_INVALID = -2,     /* bad argument (null pointer, n<0, unknown variable ...)   */
    OG_ERR_CUDA = -3,        /* a CUDA call or kernel failed; see og_last_error()        */
    OG_ERR_ALLOC = -4,       /* host or device allocation failed                         */
    OG_ERR_RADIUS = -5,      /* surface scan-1 radius below 4.5 cells: the reference     */
                             /* STOPs here (src/proc_oa.F90:694-703)                     */
    OG_ERR_NOCONV = -6       /* buddy difmax fix-point did not converge (cannot happen   */
                             /* for a finite slab; defensive)                            */
} og_status;

/* Variable codes.  OA uses 1..6 (src/proc_oa.F90:204-205), QC uses 1..7
 * (src/proc_qc.F90:264-320); code 6 is PSFC in the OA loop and PMSLPSFC in QC. */
enum {
    OG_VAR_UU = 1, OG_VAR_VV = 2, OG_VAR_TT = 3, OG_VAR_RH = 4, OG_VAR_PMSL = 5,
    OG_VAR_PSFC = 6, OG_VAR_PMSLPSFC = 6, OG_VAR_DEWPOINT = 7
};

/* QC flag bits (src/missing.inc:21-43). */
#define OG_NO_BUDDIES        (1 << 14)
#define OG_FAILS_ERROR_MAX   (1 << 16)
#define OG_FAILS_BUDDY_CHECK (1 << 17)
#define OG_OUTSIDE_OF_DOMAIN (1 << 18)
#define OG_MISSING_R         (-888888.0f)   /* src/missing.inc:14 */
#define OG_MISSING_I         (-888888)
#define OG_MAX_SCAN 10                       /* src/proc_oa.F90:105 */

typedef struct og_ctx og_ctx;

/* ---- context -------------------------------------------------------------------- */
int  og_abi_version(void);
/* device: CUDA ordinal.  Fails (OG_ERR_NO_DEVICE) unless the device is compute 10.x. */
int  og_init(og_ctx** ctx, int device);
int  og_finalize(og_ctx* ctx);
const char* og_last_error(void);
/* Counters accumulated since og_init / og_reset_stats: the library's own kernel launches,
 * Cressman contributing updates (pairs with d2 < R2, oa_module.F90:553/616), Cressman
 * candidate pairs examined, buddy pair tests examined, obs buddy-checked. */
typedef struct og_stats {
    int64_t kernel_launches;
    int64_t cressman_updates;
    int64_t cressman_candidates;
    int64_t buddy_pair_tests;
    int64_t buddy_obs;
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t buddy_fixpoint_iters;
    int64_t buddy_pairs_examined;   /* pair tests the binned kernel actually ran; buddy_pair_tests
                                       is the reference's N*(N-1) for the same slabs */
    int64_t cressman_updates_ellipse;   /* of cressman_updates: pairs of an ellipse-shaped ob     */
    int64_t cressman_updates_banana;    /* ... of a banana-shaped ob (oa_module.F90:505-542)      */
} og_stats;
int  og_get_stats(og_ctx* ctx, og_stats* out);
int  og_reset_stats(og_ctx* ctx);
/* Arithmetic of the Cressman scan kernels (the QC kernels are always exact: flags are compared
 * bit for bit).
 *   OG_ARITH_EXACT       every FP32 operation of oa_module.F90:493-620 as the reference's gfortran
 *                        build performs it (no FMA contraction, IEEE division, source order):
 *                        analysed fields are bit-identical to the reference arithmetic wherever
 *                        no banana-shaped observation contributes.
 *   OG_ARITH_CONTRACTED  the same formulas with FMA contraction, the weight written as
 *                        2 R2/(R2+d2) - 1 over the hardware reciprocal and per-observation
 *                        reciprocals for the ellipse; the summation order is unchanged.  Fields
 *                        agree with OG_ARITH_EXACT to a few FP32 ulp of the perturbation -- far
 *                        inside the 1e-5 relative / 1e-3 absolute bar -- at ~0.6x the
 *                        instructions per (grid point, observation) pair. */
enum { OG_ARITH_EXACT = 0, OG_ARITH_CONTRACTED = 1 };
int  og_set_arithmetic(og_ctx* ctx, int mode);
int  og_get_arithmetic(og_ctx* ctx);
/* When nonzero the Cressman kernels also count updates/candidates (one extra integer
 * reduction per block); off by default so the timed path does no bookkeeping. */
int  og_set_counting(og_ctx* ctx, int on);
/* When nonzero, CUDA event pairs are recorded around the Cressman scan kernels and the buddy
 * all-pairs kernel (no host synchronisation); og_last_kernel_ms() resolves them. */
int  og_set_timing(og_ctx* ctx, int on);
/* Pinned host memory for callers that want asynchronous, full-rate PCIe copies of the arrays
 * they hand to the library (Fortran: c_f_pointer on the result).  Optional: any host pointer
 * is accepted everywhere. */
int  og_host_alloc(void** ptr, size_t bytes);
int  og_host_free(void* ptr);

/* ---- per-slab drop-ins (host arrays, blocking) -----------------------------------
 * Observations must lie inside the slab, 1 <= xob < iew and 1 <= yob < jns (they are interpolated
 * from the four surrounding points); anything else, NaN included, is OG_ERR_INVALID.  The
 * reference passes obs pre-filtered to 2 <= x <= iew-1 (proc_ob_access.F90:231-234). */

/* proc_oa.F90:517-554 / :750-778.  diff = obs_value/scale - aob (or obs_value/scale when
 * use_first_guess==0); fg_value (nullable) receives aob as in :530. */
int og_innovation(og_ctx* ctx, const float* gridded, int iew, int jns,
                  const float* xob, const float* yob, const float* obs_value, int n,
                  float scale, int use_first_guess, float* diff, float* fg_value);

/* oa_module.F90:25-32.  `diff` is the reference's `obs` dummy (innovations); obs_value and
 * lat_center are unused by the reference and therefore not passed.  radius_influence is
 * the already NINT()ed integer radius in cells; dxd_km the grid distance in km
 * (driver.F90:265).  gridded is updated in place. */
int og_cressman(og_ctx* ctx, const float* diff, const float* xob, const float* yob, int n,
                float* gridded, int iew, int jns, int crsdot, int var,
                int radius_influence, float dxd_km,
                const float* u_banana, const float* v_banana, float pressure_hpa,
                int passes, int smooth_type, int use_first_guess,
                const float* fg_value, int scale_cressman_rh_decreases);

/* One (level, variable) slab of proc_oa: innovations, every Cressman scan with the
 * surface radius scaling, the re-interpolated innovations between scans and the final
 * RH clamp.  radius_influence[OG_MAX_SCAN] ends at the first negative entry.
 * scale is 100 for PMSL else 1 (proc_oa.F90:208).  n_scans_done (nullable) returns the
 * number of scans executed. */
int og_oa_slab(og_ctx* ctx, float* gridded, int iew, int jns, int var, float pressure_hpa,
               const float* obs_value, const float* xob, const float* yob, int n,
               float scale, const int* radius_influence, float radius_influence_sfc_mult,
               float dxd_km, const float* u_banana, const float* v_banana,
               int passes, int smooth_type, int use_first_guess,
               int scale_cressman_rh_decreases, int* n_scans_done);

/* qc0_module.F90:236-267: local_time = mod(int(hh + lon/15) [+24 if <0], 24) * 100.
 * time_hhmm is what proc_qc passes (time/100, proc_qc.F90:344). */
int og_local_time(og_ctx* ctx, int time_hhmm, const float* lonob, int n, int* local_time);

/* qc1_module.F90:9-15.  station_id / ship_report / index only feed prints in the
 * reference and stay on the Fortran side.  Diagnostics (nullable): aob, err, thr. */
int og_error_max(og_ctx* ctx, const float* obs, const float* xob, const float* yob,
                 const float* station_elevation, int* qc_flag, int n,
                 const float* gridded, int iew, int jns, int var, float max_difference,
                 float pressure_hpa, const int* local_time,
                 float* aob, float* err, float* thr);

/* qc2_module.F90:9-14.  Diagnostics (nullable): aob, err, difmax (after the x1.2
 * two-buddy relaxation, :327), buddy_num_final. */
int og_buddy_check(og_ctx* ctx, const float* obs, int n, const float* xob, const float* yob,
                   int* qc_flag, const float* station_elevation,
                   const float* gridded, int iew, int jns, int var, float pressure_hpa,
                   float dx_km, float buddy_weight, float buddy_difference,
                   float* aob, float* err, float* difmax, int* buddy_num_final);

/* GET_FG_T_AT_POINT (get_fg_at_point_module.F90:7-116, reached from query_ob,
 * ob_access_module.F90:257, when qc_psfc is on), batched over n points already projected to
 * grid coordinates (latlon_to_ij stays on the caller's side): first-guess temperature at
 * (x, y, height_msl) from the 3-D first-guess temperature and height fields
 * ((jns,iew,kbu) y fastest), OG_MISSING_R where the reference returns missing_r. */
int og_fg_t_at_point(og_ctx* ctx, const float* fg_3d_t, const float* fg_3d_h, int jns, int iew,
                     int kbu, const float* x_location, const float* y_location,
                     const float* height_msl, int n, float* fg_t);

/* ob_density (qc0_module.F90:81-119): tobbox(j,i) += 1 for every ob within 250 km
 * (250./grid_dist cells) of grid point (i,j); tobbox is (jns,iew), j fastest, in/out --
 * proc_qc.F90:350 accumulates it over the surface variables.
 * obs_distance (qc0_module.F90:123-232): distance (same layout, same unit as dx) from every
 * grid box to the nearest box with tobbox > 0.5; an exact Euclidean distance transform
 * instead of the reference's O((iew*jns)^2) search.  Grids up to 4096 points per edge (beyond
 * that the squared index distances are no longer exact in REAL(4): OG_ERR_INVALID). */
int og_ob_density(og_ctx* ctx, const float* xob, const float* yob, float grid_dist, int numobs,
                  float* tobbox, int iew, int jns);
int og_obs_distance(og_ctx* ctx, const float* tobbox, float* distance, int iew, int jns, float dx);

/* ---- post-analysis epilogue (host arrays, blocking) ------------------
 * 3-D arrays are the reference's (jns,iew,kbu), j fastest (first_guess.inc:1-11).
 *
 * write_analysis, final_analysis_module.F90:316-319 and :336-339, with crs2dot_pertU / _pertV
 * (first_guess_module.F90:59-91): the increment the analysis made on the A grid (u - uA) is
 * averaged onto the staggered points and added to the incoming C-grid wind; u and v are
 * replaced by the result (in/out).  Either triple may be NULL.  Bit-identical FP32. */
int og_wind_increment_to_c_grid(og_ctx* ctx, float* u, const float* uA, const float* uC,
                                float* v, const float* vA, const float* vC,
                                int jns, int iew, int kbu);
/* temporal_interp, tinterp_module.F90:53-215 (surface FDDA first guess between two analysed
 * times): out = (w1*first + w2*second)/(w1+w2) with the integer weights w1 = icount_2 -
 * icount_fdda, w2 = icount_fdda - icount_1 (both zero: w1 = 1, :106-109; sum zero:
 * OG_ERR_INVALID, the reference STOPs :112-115).  cross != 0 (TT, GHT, RH, PRES, PMSL, PMSL_C,
 * SNOW): computed on (1:jns-1,1:iew-1), last row and column replicated (:131-132); cross == 0
 * (UU, VV, UA, VA, UC, VC): the whole array.  nlev = 1 for the 2-D fields.  Bit-identical FP32. */
int og_temporal_interp(og_ctx* ctx, const float* first, const float* second, float* out,
                       int jns, int iew, int nlev, int icount_fdda, int icount_1, int icount_2,
                       int cross);
/* final_analysis_module.F90:364-370: out = pres + (pres/slp_C)*(slp_x - slp_C) for the surface
 * level of PRES (skipped by the caller when oa_psfc).  (jns,iew) arrays.  Bit-identical FP32. */
int og_pres_sfc_from_slp(og_ctx* ctx, const float* pres_sfc, const float* slp_C,
                         const float* slp_x, float* out, int jns, int iew);
/* mixing_ratio + sfcprs, final_analysis_module.F90:1026-1201 (called from proc_oa.F90:346):
 * clamps rh to [1,100] in place on (1:jns-1,1:iew-1,:), fills qv on those points of every
 * level and psfc (Pa; slp_x in Pa) on the same points of the surface; other entries of qv /
 * psfc keep the caller's values.  OG_ERR_INVALID when `pressure` lacks the 850, 700 or 500 hPa
 * level (the reference STOPs, :1122-1127).  EXP, LOG and the power operator are evaluated in FP64 and rounded once:
 * results agree with the glibc-based reference to ~1 ulp, not bit for bit. */
int og_mixing_ratio(og_ctx* ctx, float* rh, const float* t, const float* h, const float* terrain,
                    const float* slp_x, const float* pressure, int iew, int jns, int kbu,
                    float* qv, float* psfc);

/* error_max followed by buddy_check on one slab in one upload (proc_qc.F90:371-392). */
int og_qc_slab(og_ctx* ctx, int do_error_max, int do_buddy, const float* gridded, int iew,
               int jns, int var, float pressure_hpa, const float* obs, const float* xob,
               const float* yob, const float* station_elevation, const int* local_time,
               int n, float max_error, float max_buddy, float buddy_weight, float dx_km,
               int* qc_flag);

/* ---- per-time batched API -------------------------------------------------------- */

typedef struct og_qc_params {       /* namelist &record4 (src/namelist.nml) */
    int   qc_test_error_max, qc_test_buddy;
    float max_error_t, max_error_uv, max_error_rh, max_error_dewpoint, max_error_p;
    float max_buddy_t, max_buddy_uv, max_buddy_rh, max_buddy_dewpoint, max_buddy_p;
    float buddy_weight;             /* driver.F90:456 passes the literal 1. */
    int   time_hhmmss;              /* current_time_6 */
    int   qc_dewpoint;              /* run variable 7 (always on in the reference) */
    int   var_mask;                 /* 0 = every variable; else bit (v-1) enables variable v.
                                       Lets the multi-GPU driver shard one level by variable;
                                       keep RH (4) and DEWPOINT (7) on the same rank.          */
} og_qc_params;

typedef struct og_oa_params {       /* namelist &record7/8/9 */
    int   radius_influence[OG_MAX_SCAN];
    float radius_influence_sfc_mult;
    int   use_first_guess;
    int   scale_cressman_rh_decreases;
    int   smooth_type;
    int   smooth_sfc_wind, smooth_sfc_temp, smooth_sfc_rh, smooth_sfc_slp;
    int   smooth_upper_wind, smooth_upper_temp, smooth_upper_rh;
    int   surface_only;             /* fdda_loop > 1: only kp == 1 (proc_oa.F90:365) */
    int   clamp_fg_rh;              /* apply mixing_ratio's in-place RH clamp to [1,100]
                                       (final_analysis_module.F90:1055) before the loop */
    int   var_mask;                 /* 0 = every variable; else bit (v-1) enables variable v */
} og_oa_params;

/* One analysis time.  Observation table: nrep reports in the reference's report order;
 * per level k (0-based, k==0 is the 1001 hPa surface pseudo-level) and report r the
 * value at [k*nrep + r], OG_MISSING_R when the report has no such level/variable.
 * QC flags are in/out.  PMSL exists only at k==0: ob_slp / qc_slp are [nrep]. */
typedef struct og_time_desc {
    int   iew, jns, kbu;
    float dx_km;
    const float* pressure;          /* [kbu] hPa, pressure[0] == 1001 */
    float *t, *u, *v, *rh;          /* (jns,iew,kbu) y fastest; in (QC) / in-out (OA) */
    float *slp;                     /* (jns,iew) hPa */
    int   nrep;
    const float *xob, *yob;         /* [nrep] */
    const float *lon, *elev;        /* [nrep] (QC only) */
    const float *ob_u, *ob_v, *ob_t, *ob_rh;   /* [kbu*nrep] */
    const float *ob_slp;            /* [nrep] Pa */
    int *qc_u, *qc_v, *qc_t, *qc_rh;           /* [kbu*nrep] */
    int *qc_slp;                    /* [nrep] */
    /* Optional rows (NULL: absent), ABI version 3: what qc_consistency's third rule and the 'get' of
     * bogus reports need from the report list (include/obsgrid_reports.h fills them).
     *   ob_td / qc_td  [kbu*nrep] the measurements' dew point and its flag.  With both present QC
     *                  (a) stores the DEWPOINT slab's derived dew point and output flag there, as
     *                  store_ob does (ob_access_module.F90:645-657), and (b) applies the whole of
     *                  qc_consistency (qc0_module.F90:371-424): T's failed checks are copied to the RH
     *                  *and* the dew-point flag, and a dew point above the temperature fails RH and
     *                  dew point (:408-422, compared as stored, missing values included).
     *   bogus          [nrep] info%bogus: QC skips these reports (proc_ob_access.F90:179-181), OA
     *                  uses them. */
    float *ob_td;
    int   *qc_td;
    const int *bogus;
} og_time_desc;

/* proc_qc level x variable loop + qc_consistency.  Flags are updated in place. */
int og_qc_time(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p);
/* Device-resident stages over an explicit list of (level, variable mask) items instead of a level
 * range: levels strictly increasing, var_masks[i] as og_*_params.var_mask (0 / NULL: every
 * variable).  One batch -- one set of launches -- for "some slabs of the surface level plus a run of
 * upper levels", the share of a rank under level x variable sharding.  run_consistency applies to
 * the items that check every variable. */
int og_dev_qc_time_items(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p,
                         const int* levels, const int* var_masks, int nitems, int run_consistency);
int og_dev_oa_time_items(og_ctx* ctx, const og_time_desc* d, const og_oa_params* p,
                         const int* levels, const int* var_masks, int nitems);
/* qc_consistency (qc0_module.F90:371-404) of the levels [k_begin,k_end) on its own: u/v flags
 * merged, T's error_max / buddy bits copied to RH; with the optional dew-point rows of the table
 * (ob_td / qc_td) also the third rule, :405-424 -- a dew point above the temperature fails RH and
 * dew point -- and the dew-point copies of the flags (the speed / direction copies of the wind flag
 * are made by og_table_to_reports).  og_qc_time* run it themselves when
 * run_consistency != 0; the stand-alone form is the step between the QC and the OA phase when
 * the variables of a level were checked on different GPUs (var_mask; obsgrid_b200/shard.py).
 * Host arrays (blocking) and device-resident form. */
int og_qc_consistency(og_ctx* ctx, const og_time_desc* d, int k_begin, int k_end);
int og_dev_qc_consistency(og_ctx* ctx, const og_time_desc* d, int k_begin, int k_end);
/* proc_qc of a surface-FDDA time (proc_qc.F90:243-420): og_qc_time plus the observation
 * density.  tobbox (jns,iew; in/out) receives +1 per surface ob of every checked variable within
 * 250 km (ob_density at kp = 1, :349-351), is divided by 5 and zeroed below 1 (:416-417); odis
 * (jns,iew; out) is obs_distance of the result (:420), in the unit of d->dx_km. */
int og_qc_time_fdda(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p, float* tobbox,
                    float* odis);
/* proc_oa level x variable loop.  Fields are updated in place.  Observations whose flag is
 * >= min(fails_error_max, fails_buddy_check) are not used (proc_oa.F90:462).
 * k_begin/k_end select a half-open level range [k_begin,k_end) -- the unit by which the
 * multi-GPU driver shards a time across ranks; pass 0,kbu for all. */
int og_oa_time(og_ctx* ctx, const og_time_desc* d, const og_oa_params* p,
               int k_begin, int k_end);
int og_qc_time_levels(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p,
                      int k_begin, int k_end, int run_consistency);
/* proc_qc followed by proc_oa of one time period in one call (driver.F90 calls them back to
 * back): levels [k_begin,k_end) go up once, in `pipeline_groups` groups of levels (<= 0: the
 * library chooses); a group's QC + OA runs while the next group is on its way in and the
 * previous group's analysed fields and flags are on their way out, on separate copy streams.
 * Levels are independent, so the results are those of og_qc_time_levels followed
 * by og_oa_time.  Flags and fields are updated in place. */
int og_qc_oa_time(og_ctx* ctx, const og_time_desc* d, const og_qc_params* qp,
                  const og_oa_params* op, int k_begin, int k_end, int pipeline_groups);

/* ---- asynchronous per-time API: several time periods in flight ------------------------ */
/* A run over many analysis times (cfg3: 16, cfg4: 24) keeps the copy engines and the SMs busy
 * at once: while time t computes, time t+1 is on its way in and time t-1 on its way out.
 * Each of the OG_SLOTS slots owns a device mirror of one time period.
 *   og_time_upload   queues the H2D copies of levels [k_begin,k_end) of `d` (returns at once;
 *                    waits, on the device, for the slot's previous download)
 *   og_time_compute  proc_qc + proc_oa on the slot's mirror (the host arrays are not touched)
 *   og_time_download queues the D2H copies of the analysed fields and flags into `d_out`
 *                    (may be `d` itself for in-place semantics, or separate arrays)
 *   og_time_wait     blocks until the slot's download has landed
 * Host arrays should be pinned (og_host_alloc) for the copies to overlap. */
#define OG_SLOTS 4
int og_time_upload(og_ctx* ctx, const og_time_desc* d, int k_begin, int k_end, int slot);
int og_time_compute(og_ctx* ctx, const og_time_desc* d, const og_qc_params* qp,
                    const og_oa_params* op, int k_begin, int k_end, int slot);
int og_time_download(og_ctx* ctx, const og_time_desc* d_out, int k_begin, int k_end, int slot);
int og_time_wait(og_ctx* ctx, int slot);

/* The observation rows of a time period in sparse form.  Upper-air levels hold a value for the
 * soundings only (a tenth of the reports): [kbu * nrep] rows are ~90 % OG_MISSING_R there, and on a
 * box whose eight GPUs share the host's PCIe / memory fabric those bytes are what limits a run over
 * many analysis times.  Per level k the entries [lev_start[k], lev_start[k+1]) name, in ascending
 * report index, the reports that have a measurement at that level (og_reports_to_table's
 * meas_index >= 0); an absent entry means every variable missing and every flag 0.  The dense rows
 * of the og_time_desc passed alongside (ob_u .. qc_rh, ob_td, qc_td) are then ignored and may be NULL;
 * ob_slp / qc_slp, the per-report arrays and the fields are used as usual. */
typedef struct og_obs_csr {
    int   nnz;
    const int* lev_start;               /* [kbu + 1]                                   */
    const int* rep;                     /* [nnz] report index, 0-based                 */
    const float *u, *v, *t, *rh, *td;   /* [nnz]; td nullable (with qc_td)             */
    int *qc_u, *qc_v, *qc_t, *qc_rh, *qc_td;   /* [nnz] in (upload) / out (download)   */
} og_obs_csr;
/* og_time_upload / og_time_download with the observation rows in that form: the upload expands them
 * into the slot's dense device mirror, the download packs the flag rows (and the dew points the
 * DEWPOINT slabs derived, when td is present: obs->td must then be writable) back into `obs`.
 * og_time_compute and og_time_wait are used unchanged. */
int og_time_upload_csr(og_ctx* ctx, const og_time_desc* d, const og_obs_csr* obs, int k_begin, int k_end, int slot);
int og_time_download_csr(og_ctx* ctx, const og_time_desc* d_out, og_obs_csr* obs_out, int k_begin, int k_end, int slot);

/* ---- resident (device-pointer) API ---------------------------------------------- */
/* Same semantics as og_qc_time / og_oa_time but every pointer inside `d` is a DEVICE
 * pointer on the context's device and nothing is copied; used to time the kernels with
 * inputs already in HBM.  `pressure` stays a host pointer. */
int og_dev_qc_time(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p,
                   int k_begin, int k_end, int run_consistency);
int og_dev_oa_time(og_ctx* ctx, const og_time_desc* d, const og_oa_params* p,
                   int k_begin, int k_end);
/* ---- one (level, variable) slab over several GPUs ----------------------------------------------
 * The buddy check of a dense surface network is the largest indivisible unit of a time period (2 M
 * reports: tens of milliseconds), so a run that strong-scales one analysis time has to cut inside the
 * slab.  With an exchange function installed, og_*qc_time* on a context of rank r of `world`:
 *   - every rank prepares every observation of the batch (error_max, difmax: O(N));
 *   - the buddy check's cells are dealt to the ranks (cell index mod world): a rank builds the
 *     candidate lists of its own cells only and runs the pair kernels for the observations sitting
 *     in them -- each target's count and index-ordered sum are computed by exactly one rank, with
 *     the very operations of the unsharded run;
 *   - the relaxation mask of qc2_module.F90:327 is combined after every pass of the fix-point
 *     (exchange, OG_XCHG_MAX on bytes), the flags once at the end (OG_XCHG_MAX on int32: only the
 *     owner of an observation adds bits to the common starting value, so MAX is the bitwise OR).
 * and og_*oa_time*, og_oa_slab, og_cressman:
 *   - the 32x32-point cells of every slab are dealt the same way (cell index mod world): a rank
 *     builds the lists of its own cells and scans them; other cells are written as zero bits, and
 *     after every scan the slab (or, on the smoothing path, the perturbation) is summed over the
 *     ranks as int32 (OG_XCHG_SUM: x + 0 + ... + 0 is exact), so that the innovations of the next
 *     scan (proc_oa.F90:751-763) are interpolated from the complete field on every rank:
 *     iew * jns * 4 bytes per slab and scan.
 * The result on every rank is bit-identical to the single-GPU result.  Worth it for the few big
 * slabs that whole-slab dealing cannot balance (the surface level); a batch of many small slabs is
 * better dealt slab by slab (one exchange call per slab and scan is issued here).
 *
 * exchange(user, buf, count, dtype, op, stream): all-reduce `count` elements at device pointer `buf`
 * in place over the ranks, ordered on the CUDA stream `stream` (a cudaStream_t); must return 0.
 * NCCL: ncclAllReduce(buf, buf, count, dtype, op, comm, stream); obsgrid_b200/shard.py wraps
 * torch.distributed. */
enum { OG_XCHG_INT32 = 0, OG_XCHG_UINT8 = 1 };
enum { OG_XCHG_SUM = 0, OG_XCHG_MAX = 1 };
typedef int (*og_exchange_fn)(void* user, void* buf, size_t count, int dtype, int op, void* stream);
/* world <= 1 or fn == NULL: back to single-GPU behaviour. */
int og_set_exchange(og_ctx* ctx, int rank, int world, og_exchange_fn fn, void* user);

/* Device-resident forms of og_temporal_interp and og_qc_time_fdda (every array a device pointer):
 * with og_dev_qc_time / og_dev_oa_time they compose the surface-FDDA chain of driver.F90:107-113,
 * :306-341, :561-563 without a field leaving HBM between its phases -- the 3-D analyses of the
 * bracketing times, store_fa, temporal_interp to every FDDA time, proc_qc with the observation
 * density and the surface-only proc_oa (obsgrid_b200/fdda.py composes them). */
int og_dev_temporal_interp(og_ctx* ctx, const float* first, const float* second, float* out,
                           int jns, int iew, int nlev, int icount_fdda, int icount_1, int icount_2,
                           int cross);
int og_dev_qc_time_fdda(og_ctx* ctx, const og_time_desc* d, const og_qc_params* p, float* tobbox,
                        float* odis);
/* field(j,i) *= factor on a (jns,iew) device array; interior_only: (1:jns-1, 1:iew-1) only -- store_fa
 * keeps PMSL in Pa (slp_x * 100., tinterp_module.F90:45) and the driver brings the interpolated field
 * back with slp_x(:jns-1,:iew-1) * 0.01 (driver.F90:322, :336). */
int og_dev_scale_field(og_ctx* ctx, float* field, int jns, int iew, float factor, int interior_only);
/* Stream all library work is issued on (a cudaStream_t), for event timing by the caller. */
void* og_stream(og_ctx* ctx);
int   og_sync(og_ctx* ctx);
/* Kernel-only time accumulated since the last call (og_set_timing must be on), measured with
 * CUDA events on the library's stream: ms inside the Cressman scan kernels and inside the
 * buddy all-pairs kernel.  Synchronises the stream. */
int   og_last_kernel_ms(og_ctx* ctx, float* cressman_ms, float* buddy_ms);
/* Micro-benchmark: sustained FP32 lane-instructions/s of dependent-chain FADD/FMUL
 * (no FMA contraction) on this device -- the denominator of the FP32 roofline. */
int   og_fp32_peak(og_ctx* ctx, double* lane_ops_per_s);
/* Self test: compares the kernels' FCHK-less exact division with __fdiv_rn on n random
 * Cressman weight operand pairs (R in 1..rmax); *mismatches must come back 0. */
int   og_selftest_div(og_ctx* ctx, int n, unsigned seed, int rmax, long long* mismatches);
/* Self test: out[i] = the device's LOG of x[i] (host arrays).  The DEWPOINT quantities of proc_qc go
 * through Fortran's LOG, i.e. glibc's logf in the reference build (proc_qc.F90:315-319,
 * ob_access_module.F90:476-486); the device restates that algorithm (csrc/og_logf.h) and the tests
 * compare it with the host libm's logf over the whole argument range. */
int   og_selftest_logf(og_ctx* ctx, const float* x, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* OBSGRID_GPU_H */
