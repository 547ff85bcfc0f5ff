/*
 * og_oracle.h -- CPU oracle for the OBSGRID objective-analysis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the CPU baseline.  libobsgrid_gpu.so never
 * links or calls it.
 *
 * PARITY UNPINNED: the reference (wrf-model/OBSGRID) ships no tests, golden vectors or
 * fixtures, and is Fortran 90 + netCDF, neither of which exists in the build container
 * (no gfortran/ifort/nvfortran/flang, no netcdf.inc), so oracle/_ref cannot be built.
 * This file is a line-by-line FP32 restatement of the reference loops; it is pinned only
 * by the hand-derived known-answer tests in tests/test_oracle_kat.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile) so that every
 * operation is one IEEE-754 binary32 operation in source order, as gfortran -O emits on
 * x86-64 (arch/configure.defaults:44-46: no fast-math, no FMA contraction).
 */
#ifndef OG_ORACLE_H
#define OG_ORACLE_H

#include <stdint.h>
#include "../include/obsgrid_gpu.h"   /* shared plain-C structs/constants only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ogo_stats {
    int64_t cressman_updates;      /* executions of oa_module.F90:566-568 / :618-620 */
    int64_t cressman_candidates;   /* inner-loop iterations (window sweep)           */
    int64_t buddy_pair_tests;      /* executions of qc2_module.F90:233-236           */
    int64_t n_circle, n_ellipse, n_banana;
    int64_t buddy_pairs_examined;  /* binned fast path: pair tests actually run             */
} ogo_stats;

void ogo_reset_stats(void);
/* Fast paths for the large configurations (same FP32 operations, re-scheduled; bit-identical to
 * the as-written loops, tests/test_oracle_fast_paths.py).  buddy_binned: spatial bins instead of
 * the N^2 station loops of buddy_check; inner_threads: pthreads inside one cressman / buddy_check
 * call (grid rows / target bins dealt out). */
void ogo_set_fast_paths(int buddy_binned, int inner_threads);
void ogo_get_stats(ogo_stats* out);

void ogo_innovation(const float* gridded, int iew, int jns, const float* xob,
                    const float* yob, const float* obs_value, int n, float scale,
                    int use_first_guess, float* diff, float* fg_value);

void ogo_cressman(const float* diff, const float* xob, const float* yob, int n,
                  float* gridded, int iew, int jns, int crsdot, int var,
                  int radius_influence, float dxd_km, const float* u_banana,
                  const float* v_banana, float pressure_hpa, int passes, int smooth_type,
                  int use_first_guess, const float* fg_value,
                  int scale_cressman_rh_decreases);

int ogo_oa_slab(float* gridded, int iew, int jns, int var, float pressure_hpa,
                const float* obs_value, const float* xob, const float* yob, int n,
                float scale, const int* radius_influence, float radius_influence_sfc_mult,
                float dxd_km, const float* u_banana, const float* v_banana, int passes,
                int smooth_type, int use_first_guess, int scale_cressman_rh_decreases,
                int* n_scans_done);

void ogo_smooth_5(float* field, int iew, int jns, int passes, int crsdot);
void ogo_smoother_desmoother(float* slab, int imx, int jmx, int passes, int crsdot);
void ogo_clean_rh(float* rh, int iew, int jns, float rh_min, float rh_max);

void ogo_modify_qc_flag(int* qc_flag, int n, const float* err, const float* thr, int test);
void ogo_add_to_qc_flag(int* qc_flag, int to_add);
int ogo_contains_2n(int sumtocheck, int numtofind);        /* obs_sort_module.F90:4515-4560 */
void ogo_local_time(int time_hhmm, const float* lonob, int n, int* local_time);

void ogo_error_max(const float* obs, const float* xob, const float* yob,
                   const float* station_elevation, int* qc_flag, int n,
                   const float* gridded, int iew, int jns, int var, float max_difference,
                   float pressure_hpa, const int* local_time,
                   float* aob, float* err, float* thr);

void ogo_buddy_check(const float* obs, int n, const float* xob, const float* yob,
                     int* qc_flag, const float* station_elevation, const float* gridded,
                     int iew, int jns, int var, float pressure_hpa, float dx_km,
                     float buddy_weight, float buddy_difference,
                     float* aob, float* err, float* difmax, int* buddy_num_final);

/* ob_density / obs_distance (qc0_module.F90:81-232): observation density and distance to the
 * nearest observed grid box for the surface FDDA.  tobbox / distance are (jns,iew), j fastest. */
void ogo_ob_density(const float* xob, const float* yob, float grid_dist, int numobs,
                    float* tobbox, int iew, int jns);
void ogo_obs_distance(const float* tobbox, float* distance, int iew, int jns, float dx);

/* DEWPOINT first-guess field of proc_qc.F90:315-319 and dewpoint ob of
 * ob_access_module.F90:476-486 (same formula). */
float ogo_dewpoint(float t, float rh);
/* libm's logf (what gfortran's LOG calls) over an array. */
void ogo_logf_array(const float* x, int n, float* out);

/* proc_qc / proc_oa level x variable loops over the flat observation table. */
int ogo_qc_time(const og_time_desc* d, const og_qc_params* p, int k_begin, int k_end,
                int run_consistency);
int ogo_oa_time(const og_time_desc* d, const og_oa_params* p, int k_begin, int k_end);
/* qc_consistency (qc0_module.F90:371-404) of the levels [k_begin,k_end) on its own. */
void ogo_qc_consistency(const og_time_desc* d, int k_begin, int k_end);
/* proc_qc of a surface-FDDA time: ogo_qc_time over every level + ob_density at kp = 1,
 * tobbox / 5 thresholded at 1, obs_distance (proc_qc.F90:349-351, :412-420). */
int ogo_qc_time_fdda(const og_time_desc* d, const og_qc_params* p, float* tobbox, float* odis);
/* Same as ogo_oa_time but slabs are distributed over `threads` pthreads (slabs of one
 * level are independent) -- the "all host cores" CPU baseline. */
int ogo_oa_time_mt(const og_time_desc* d, const og_oa_params* p, int k_begin, int k_end,
                   int threads);
int ogo_qc_time_mt(const og_time_desc* d, const og_qc_params* p, int k_begin, int k_end,
                   int run_consistency, int threads);

/* The reference's observation access as written, over the plain-array image of the report list
 * (og_oracle_reports.c; structs in ../include/obsgrid_reports.h): query_ob, store_ob, proc_ob_access
 * 'get' / 'use', qc_consistency over every measurement, and proc_qc's loop that calls them. */
struct og_report; struct og_meas;
void ogo_query_ob(struct og_report* obs, struct og_meas* meas, int date, int time, int var, float request_level,
                  int request_qc_max, int request_p_diff, float* value, int* qc);
int ogo_store_ob(struct og_report* obs, struct og_meas* meas, int var, float request_level, int request_p_diff,
                 float value, int qc);
int ogo_get_obs(struct og_report* obs, int total_numobs, struct og_meas* meas, int is_get, int var,
                float request_level, int date, int time, int request_p_diff, int request_qc_max, int iew_alloc,
                int jns_alloc, float* get_value, float* get_x, float* get_y, float* get_lon, float* get_elev,
                int* get_index, int* get_qc);
void ogo_qc_consistency_reports(struct og_report* obs, int num_obs, struct og_meas* meas);
int ogo_qc_time_reports(const og_time_desc* d, const og_qc_params* p, struct og_report* obs, int nrep,
                        struct og_meas* meas, int date, int time, int request_p_diff);

/* GET_FG_T_AT_POINT (get_fg_at_point_module.F90:7-116) with (x,y) already projected. */
float ogo_fg_t_at_point(const float* t3d, const float* h3d, int jns, int iew, int kbu,
                        float x_location, float y_location, float height_msl);

/* Post-analysis epilogue.  3-D arrays are (jns,iew,kbu), j fastest.
 * final_analysis_module.F90:316-350 + first_guess_module.F90:59-91 (A-grid increments onto the
 * C-grid winds), :364-370 (surface PRES from the SLP increment), :1026-1201 (mixing_ratio +
 * sfcprs; expf/logf/powf are glibc's, as in the gfortran build). */
void ogo_wind_increment_u(float* u, const float* uA, const float* uC, int jns, int iew, int kbu);
void ogo_wind_increment_v(float* v, const float* vA, const float* vC, int jns, int iew, int kbu);
void ogo_pres_sfc_from_slp(const float* pres_sfc, const float* slp_C, const float* slp_x,
                           float* out, int jns, int iew);
int ogo_temporal_interp(const float* first, const float* second, float* out, int jns, int iew,
                        int nlev, int icount_fdda, int icount_1, int icount_2, int cross);
int ogo_mixing_ratio(float* rh, const float* t, const float* h, const float* terrain,
                     const float* slp_x, const float* pressure, int iew, int jns, int kbu,
                     float* qv, float* psfc);

#ifdef __cplusplus
}
#endif
#endif
