/*
 * og_oracle.c -- CPU oracle: FP32 restatement of OBSGRID's objective-analysis hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see og_oracle.h).  PARITY UNPINNED by the reference: it has
 * no tests/golden vectors and cannot be compiled here (Fortran + netCDF absent); pinned by
 * the hand-derived known-answer tests in tests/test_oracle_kat.py.
 *
 * Every function cites the reference lines it restates (paths relative to the reference
 * tree).  Fortran semantics kept: REAL = binary32, literals single precision, INT() and
 * real->integer assignment truncate toward zero, NINT rounds half away from zero,
 * x**2 = x*x, x**3 = (x*x)*x, SIGN(1.,v) = copysignf(1,v), strict left-to-right
 * evaluation with the source's parentheses, no FMA contraction (compile with
 * -ffp-contract=off), arrays 1-based and column-major.
 */
#define _GNU_SOURCE
#include "og_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------- */
static ogo_stats g_stats;

void ogo_reset_stats(void) { memset(&g_stats, 0, sizeof g_stats); }
void ogo_get_stats(ogo_stats* out) { *out = g_stats; }
static void stats_add(int64_t* dst, int64_t v) { __atomic_fetch_add(dst, v, __ATOMIC_RELAXED); }

/* Fast paths for the large configurations (BASELINE configs 3-5: up to 2 M reports on 4000^2 points).
 * The as-written loops above/below stay the definition; these re-schedule the very same FP32
 * operations and are held to them bit for bit by tests/test_oracle_fast_paths.py:
 *   inner_threads > 1  ogo_cressman deals the grid rows to pthreads (a grid point's sums see the obs
 *                      in the same ascending order); the binned buddy check deals the target bins.
 *   buddy_binned != 0  the station loops of buddy_check visit spatial bins one buddy range wide
 *                      instead of all N^2 pairs. */
static int g_inner_threads = 1, g_buddy_binned = 0;
void ogo_set_fast_paths(int buddy_binned, int inner_threads)
{
    g_buddy_binned = buddy_binned;
    g_inner_threads = inner_threads > 1 ? inner_threads : 1;
}

static inline int f_nint(float x) { return (int)lroundf(x); } /* NINT */
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* 1-based column-major access to an (iew,jns) slab */
#define AT(a, i, j) ((a)[((size_t)((j)-1)) * (size_t)iew + (size_t)((i)-1)])

/* Bilinear first guess at an ob -- proc_oa.F90:520-527, qc1_module.F90:197-204,
 * qc2_module.F90:204-211 (identical expressions). */
static inline float bilinear(const float* g, int iew, float x, float y)
{
    int iob = (int)x;
    int job = (int)y;
    float dxob = x - (float)iob;
    float dyob = y - (float)job;
    return (1.f - dxob) * ((1.f - dyob) * AT(g, iob, job) + dyob * AT(g, iob, job + 1)) +
           dxob * ((1.f - dyob) * AT(g, iob + 1, job) + dyob * AT(g, iob + 1, job + 1));
}

/* proc_oa.F90:517-554 (scan 1) and :750-778 (after every scan). */
void ogo_innovation(const float* gridded, int iew, int jns, const float* xob,
                    const float* yob, const float* obs_value, int n, float scale,
                    int use_first_guess, float* diff, float* fg_value)
{
    (void)jns;
    for (int num = 0; num < n; ++num) {
        float aob = bilinear(gridded, iew, xob[num], yob[num]);
        if (use_first_guess) {
            diff[num] = obs_value[num] / scale - aob;
            if (fg_value) fg_value[num] = aob; /* :530 */
        } else {
            diff[num] = obs_value[num] / scale; /* :551 */
        }
    }
}

/* oa_module.F90:11-20 */
void ogo_clean_rh(float* rh, int iew, int jns, float rh_min, float rh_max)
{
    size_t n = (size_t)iew * jns;
    for (size_t i = 0; i < n; ++i) if (rh[i] > rh_max) rh[i] = rh_max;
    for (size_t i = 0; i < n; ++i) if (rh[i] < rh_min) rh[i] = rh_min;
}

/* oa_module.F90:928-1006 */
void ogo_smooth_5(float* field, int iew, int jns, int passes, int crsdot)
{
    float* temp = (float*)malloc(sizeof(float) * (size_t)iew * jns);
    for (int np = 1; np <= passes; ++np) {
        for (int j = 2; j <= jns - 1 - crsdot; ++j)
            for (int i = 2; i <= iew - 1 - crsdot; ++i)
                AT(temp, i, j) = (AT(field, i, j) * 4.f + AT(field, i + 1, j) +
                                  AT(field, i - 1, j) + AT(field, i, j + 1) +
                                  AT(field, i, j - 1)) * 0.125f;
        int i = 1;
        for (int j = 2; j <= jns - 1 - crsdot; ++j)
            AT(temp, i, j) = (AT(field, i, j) * 2.f + AT(field, i, j + 1) + AT(field, i, j - 1)) * 0.25f;
        i = iew - crsdot;
        for (int j = 2; j <= jns - 1 - crsdot; ++j)
            AT(temp, i, j) = (AT(field, i, j) * 2.f + AT(field, i, j + 1) + AT(field, i, j - 1)) * 0.25f;
        int j = 1;
        for (i = 2; i <= iew - 1 - crsdot; ++i)
            AT(temp, i, j) = (AT(field, i, j) * 2.f + AT(field, i + 1, j) + AT(field, i - 1, j)) * 0.25f;
        j = jns - crsdot;
        for (i = 2; i <= iew - 1 - crsdot; ++i)
            AT(temp, i, j) = (AT(field, i, j) * 2.f + AT(field, i + 1, j) + AT(field, i - 1, j)) * 0.25f;

        for (j = 2; j <= jns - 1 - crsdot; ++j)
            for (i = 2; i <= iew - 1 - crsdot; ++i) AT(field, i, j) = AT(temp, i, j);
        for (j = 2; j <= jns - 1 - crsdot; ++j) {
            AT(field, 1, j) = AT(temp, 1, j);
            AT(field, iew - crsdot, j) = AT(temp, iew - crsdot, j);
        }
        for (i = 2; i <= iew - 1 - crsdot; ++i) {
            AT(field, i, 1) = AT(temp, i, 1);
            AT(field, i, jns - crsdot) = AT(temp, i, jns - crsdot);
        }
    }
    free(temp);
}

/* oa_module.F90:875-923 */
void ogo_smoother_desmoother(float* slab, int imx, int jmx, int passes, int crsdot)
{
    const int iew = imx; /* for AT() */
    float* slabnew = (float*)malloc(sizeof(float) * (size_t)imx * jmx);
    const float xnu[2] = {0.50f, -0.52f};
    for (int loop = 1; loop <= passes * 2; ++loop) {
        int n = 2 - (loop % 2); /* 1-based index into xnu */
        float nu = xnu[n - 1];
        for (int i = 2; i <= imx - 1 - crsdot; ++i)
            for (int j = 2; j <= jmx - 1 - crsdot; ++j)
                AT(slabnew, i, j) = AT(slab, i, j) +
                    nu * ((AT(slab, i, j + 1) + AT(slab, i, j - 1)) * 0.5f - AT(slab, i, j));
        for (int i = 2; i <= imx - 1 - crsdot; ++i)
            for (int j = 2; j <= jmx - 1 - crsdot; ++j) AT(slab, i, j) = AT(slabnew, i, j);
        for (int j = 2; j <= jmx - 1 - crsdot; ++j)
            for (int i = 2; i <= imx - 1 - crsdot; ++i)
                AT(slabnew, i, j) = AT(slab, i, j) +
                    nu * ((AT(slab, i + 1, j) + AT(slab, i - 1, j)) * 0.5f - AT(slab, i, j));
        for (int i = 2; i <= imx - 1 - crsdot; ++i)
            for (int j = 2; j <= jmx - 1 - crsdot; ++j) AT(slab, i, j) = AT(slabnew, i, j);
    }
    free(slabnew);
}

/* ------------------------------------------------------------------------------- */
/* oa_module.F90:577-627 */
static void dist_not_uu_vv_rh(int iew_min, int iew_max, int jns_min, int jns_max, int iew,
                              float rc2, float obs_n, float xob, float yob,
                              int radius_influence, float* numerator, float* denominator,
                              int64_t* hits, int64_t* cands)
{
    float r2 = (float)(radius_influence * radius_influence);
    for (int jj = jns_min; jj <= jns_max; ++jj) {
        float real_j = (float)jj;
        for (int ii = iew_min; ii <= iew_max; ++ii) {
            float real_i = (float)ii;
            float d2 = ((xob - rc2 - real_i) * (xob - rc2 - real_i)) +
                       ((yob - rc2 - real_j) * (yob - rc2 - real_j));
            ++*cands;
            if (d2 < r2) {
                float weight = (r2 - d2) / (r2 + d2);
                AT(numerator, ii, jj) = AT(numerator, ii, jj) + weight * weight * obs_n;
                AT(denominator, ii, jj) = AT(denominator, ii, jj) + weight;
                ++*hits;
            }
        }
    }
}

/* oa_module.F90:229-574 */
static void dist_uu_vv_rh(int iew_min, int iew_max, int jns_min, int jns_max, int iew, int jns,
                          float rc2, int var, float obs_n, float xob, float yob, float dxd,
                          float pressure, const float* u_banana, const float* v_banana,
                          int radius_influence, const float* gridded, float fg_value_n,
                          int use_first_guess, float* numerator_pert, float* denominator_pert,
                          int scale_cressman_rh_decreases, int64_t* hits, int64_t* cands,
                          int* shape_out)
{
    const float pi = 3.14159265358f;
    const float twopi = 2.f * 3.14159265358f;
    int use_circle = 0, use_banana = 0, use_ellipse = 0;
    float r2 = (float)(radius_influence * radius_influence);
    float vcrit;
    float radius_of_curvature = 0.f, elongation = 0.f, theta_ob = 0.f;
    float one_divided_by_twopi = 0.f, one_divided_by_pi = 0.f;
    int i_center = 0, j_center = 0;

    if (pressure < 500.f) vcrit = 15.f;           /* :309-313 */
    else vcrit = 25.f - 0.02f * pressure;

    float u_ob = AT(u_banana, f_nint(xob), f_nint(yob)); /* :332-334 */
    float v_ob = AT(v_banana, f_nint(xob), f_nint(yob));
    float speed_ob = sqrtf(u_ob * u_ob + v_ob * v_ob);

    if (speed_ob >= vcrit) { /* :340-397 */
        float dydx;
        if (fabsf(u_ob) >= 1.E-5f) dydx = v_ob / u_ob;
        else dydx = 1000.f;
        float x1 = xob + 3.f * (u_ob / speed_ob);
        float y1 = yob + 3.f * (v_ob / speed_ob);
        float x2 = xob - 3.f * (u_ob / speed_ob);
        float y2 = yob - 3.f * (v_ob / speed_ob);
        int i1 = imax(imin(f_nint(x1), iew), 1);
        int j1 = imax(imin(f_nint(y1), jns), 1);
        int i2 = imax(imin(f_nint(x2), iew), 1);
        int j2 = imax(imin(f_nint(y2), jns), 1);
        float d12 = sqrtf(((float)i1 - (float)i2) * ((float)i1 - (float)i2) +
                          ((float)j1 - (float)j2) * ((float)j1 - (float)j2));
        float dudx = (AT(u_banana, i1, j1) - AT(u_banana, i2, j2)) / (d12 * dxd);
        float dvdx = (AT(v_banana, i1, j1) - AT(v_banana, i2, j2)) / (d12 * dxd);
        float numerator_curvature = fabsf(u_ob * dvdx - v_ob * dudx) / (u_ob * u_ob);
        float s = sqrtf(1.f + dydx * dydx);
        float denominator_curvature = (s * s) * s;
        if ((fabsf(u_ob) < 1.E-5f) || (fabsf(v_ob) < 1.E-5f) ||
            (fabsf(numerator_curvature) < 1.E-5f))
            radius_of_curvature = 1000.f;
        else
            radius_of_curvature = denominator_curvature / numerator_curvature;
    }

    if (speed_ob < vcrit) { /* :403 */
        use_circle = 1;
    } else if (fabsf(radius_of_curvature) < (3.f * (float)radius_influence) * dxd) { /* :410 */
        use_banana = 1;
        one_divided_by_twopi = 1.f / twopi;
        one_divided_by_pi = 1.f / pi;
        int ia = imax(f_nint(xob) - 3, 1);
        int ja = (int)yob;
        int ib = (int)xob;
        int jb = imax(f_nint(yob) - 3, 1);
        int ic = imin(f_nint(xob) + 3, iew);
        int jc = (int)yob;
        int id = (int)xob;
        int jd = imin(f_nint(yob) + 3, jns);
        float va = AT(v_banana, ia, ja);
        float ub = AT(u_banana, ib, jb);
        float vc = AT(v_banana, ic, jc);
        float ud = AT(u_banana, id, jd);
        float vor = (vc - va) - (ud - ub);
        radius_of_curvature = radius_of_curvature * copysignf(1.f, vor);        /* :451 */
        i_center = (int)((float)f_nint(xob) - radius_of_curvature * v_ob / speed_ob); /* :458 */
        j_center = (int)((float)f_nint(yob) + radius_of_curvature * u_ob / speed_ob);
        theta_ob = atan2f((float)j_center - yob, (float)i_center - xob);        /* :466 */
        elongation = 1.f + 0.7778f / vcrit * speed_ob;
    } else {
        use_ellipse = 1;
        elongation = 1.f + 0.7778f / vcrit * speed_ob;
    }
    if (shape_out) *shape_out = use_circle ? 0 : (use_ellipse ? 1 : 2);

    int do_rh_scaling = (scale_cressman_rh_decreases && (var == OG_VAR_RH) && use_first_guess &&
                         (obs_n < 0.0f)); /* :485-490 */

    for (int jj = jns_min; jj <= jns_max; ++jj) {
        float real_j = (float)jj;
        for (int ii = iew_min; ii <= iew_max; ++ii) {
            float real_i = (float)ii;
            float distance_squared;
            if (use_circle) {
                distance_squared = ((xob - rc2 - real_i) * (xob - rc2 - real_i)) +
                                   ((yob - rc2 - real_j) * (yob - rc2 - real_j));
            } else if (use_banana) {
                float theta_ij = atan2f((float)j_center - real_j, (float)i_center - real_i);
                float rij = sqrtf(((float)i_center - real_i) * ((float)i_center - real_i) +
                                  ((float)j_center - real_j) * ((float)j_center - real_j));
                float theta_diff = theta_ob - theta_ij;
                if (theta_diff > twopi)
                    theta_diff = theta_diff - (float)((int)(theta_diff * one_divided_by_twopi)) * twopi;
                if (theta_diff < -twopi)
                    theta_diff = theta_diff + (float)((int)(theta_diff * one_divided_by_twopi)) * twopi;
                if (theta_diff > pi) theta_diff = theta_diff - twopi;
                if (theta_diff < -pi) theta_diff = theta_diff + twopi;
                float x = radius_of_curvature * theta_diff * one_divided_by_pi;
                float y = fabsf(radius_of_curvature) - rij;
                distance_squared = x * x / elongation + y * y;
            } else {
                float x = (u_ob * (real_i - xob) + v_ob * (real_j - yob)) / speed_ob;
                float y = (v_ob * (real_i - xob) - u_ob * (real_j - yob)) / speed_ob;
                distance_squared = x * x / elongation + y * y;
            }
            ++*cands;
            if (distance_squared < r2) {
                float scale_factor = 1.0f;
                if (do_rh_scaling) scale_factor = fminf(1.0f, AT(gridded, ii, jj) / fg_value_n);
                float weight = (r2 - distance_squared) / (r2 + distance_squared);
                AT(numerator_pert, ii, jj) =
                    AT(numerator_pert, ii, jj) + weight * weight * obs_n * scale_factor;
                AT(denominator_pert, ii, jj) = AT(denominator_pert, ii, jj) + weight;
                ++*hits;
            }
        }
    }
}

/* The observation loop of cressman (oa_module.F90:116-190) restricted to the grid rows
 * [j_lo, j_hi].  One band = all rows is the loop as written; several bands on several threads
 * perform the same additions per grid point in the same (ascending ob) order. */
typedef struct {
    const float* obs; const float* xob; const float* yob; int numobs;
    const float* gridded; int iew, jns, crsdot, var, radius_influence; float dxd;
    const float* u_banana; const float* v_banana; float pressure;
    int use_first_guess; const float* fg_value; int scale_cressman_rh_decreases;
    float* numerator; float* denominator;
    int j_lo, j_hi;
    int64_t hits, cands, nshape[3];
} cressman_band;

static void* cressman_band_run(void* arg)
{
    cressman_band* B = (cressman_band*)arg;
    const float* xob = B->xob; const float* yob = B->yob;
    const int iew = B->iew, jns = B->jns, crsdot = B->crsdot, radius_influence = B->radius_influence;
    const int var = B->var;
    float rc2 = (float)crsdot / 2.f;
    for (int n = 0; n < B->numobs; ++n) {
        if ((xob[n] <= 1.f + rc2) || (yob[n] <= 1.f + rc2) ||
            (xob[n] >= (float)iew - rc2) || (yob[n] >= (float)jns - rc2))
            continue; /* :134-140 */
        int iew_min = imax(f_nint(xob[n] - rc2) - 4 * radius_influence - 1, 1);
        int jns_min = imax(f_nint(yob[n] - rc2) - 4 * radius_influence - 1, 1);
        int iew_max = imin(f_nint(xob[n] - rc2) + 4 * radius_influence + 1, iew - crsdot);
        int jns_max = imin(f_nint(yob[n] - rc2) + 4 * radius_influence + 1, jns - crsdot);
        const int first_band = (B->j_lo == 1);       /* shapes are counted once per ob */
        jns_min = imax(jns_min, B->j_lo); jns_max = imin(jns_max, B->j_hi);
        if (jns_min > jns_max && !first_band) continue;
        if (var == OG_VAR_UU || var == OG_VAR_VV || var == OG_VAR_RH) {
            int shape = 0;
            dist_uu_vv_rh(iew_min, iew_max, jns_min, jns_max, iew, jns, rc2, var, B->obs[n],
                          xob[n], yob[n], B->dxd, B->pressure, B->u_banana, B->v_banana,
                          radius_influence, B->gridded, B->fg_value ? B->fg_value[n] : 0.f,
                          B->use_first_guess, B->numerator, B->denominator,
                          B->scale_cressman_rh_decreases, &B->hits, &B->cands, &shape);
            if (first_band) B->nshape[shape]++;
        } else {
            dist_not_uu_vv_rh(iew_min, iew_max, jns_min, jns_max, iew, rc2, B->obs[n], xob[n],
                              yob[n], radius_influence, B->numerator, B->denominator, &B->hits,
                              &B->cands);
            if (first_band) B->nshape[0]++;
        }
    }
    return NULL;
}

/* oa_module.F90:25-218 */
void ogo_cressman(const float* obs, const float* xob, const float* yob, int numobs,
                  float* gridded, int iew, int jns, int crsdot, int var,
                  int radius_influence, float dxd, const float* u_banana,
                  const float* v_banana, float pressure, int passes, int smooth_type,
                  int use_first_guess, const float* fg_value,
                  int scale_cressman_rh_decreases)
{
    size_t npts = (size_t)iew * jns;
    float* numerator = (float*)calloc(npts, sizeof(float));
    float* denominator = (float*)calloc(npts, sizeof(float));
    float* perturbation = (float*)calloc(npts, sizeof(float));
    int64_t hits = 0, cands = 0, nshape[3] = {0, 0, 0};

    const int nband = (g_inner_threads > 1 && jns >= 4 * g_inner_threads) ? g_inner_threads : 1;
    cressman_band* bands = (cressman_band*)calloc((size_t)nband, sizeof(cressman_band));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nband);
    for (int b = 0; b < nband; ++b) {
        bands[b] = (cressman_band){obs, xob, yob, numobs, gridded, iew, jns, crsdot, var,
                                   radius_influence, dxd, u_banana, v_banana, pressure,
                                   use_first_guess, fg_value, scale_cressman_rh_decreases,
                                   numerator, denominator,
                                   1 + (int)((long long)jns * b / nband),
                                   (int)((long long)jns * (b + 1) / nband), 0, 0, {0, 0, 0}};
        if (nband > 1) pthread_create(&th[b], NULL, cressman_band_run, &bands[b]);
        else cressman_band_run(&bands[b]);
    }
    for (int b = 0; b < nband; ++b) {
        if (nband > 1) pthread_join(th[b], NULL);
        hits += bands[b].hits; cands += bands[b].cands;
        for (int q = 0; q < 3; ++q) nshape[q] += bands[b].nshape[q];
    }
    free(bands); free(th);
    for (size_t p = 0; p < npts; ++p) /* :195-197 */
        if (denominator[p] > 0.f) perturbation[p] = numerator[p] / denominator[p];

    if (smooth_type == 1) ogo_smooth_5(perturbation, iew, jns, passes, crsdot);
    else if (smooth_type == 2) ogo_smoother_desmoother(perturbation, iew, jns, passes, crsdot);

    if (use_first_guess)
        for (size_t p = 0; p < npts; ++p) gridded[p] = gridded[p] + perturbation[p];
    else
        for (size_t p = 0; p < npts; ++p) gridded[p] = perturbation[p];

    stats_add(&g_stats.cressman_updates, hits);
    stats_add(&g_stats.cressman_candidates, cands);
    stats_add(&g_stats.n_circle, nshape[0]);
    stats_add(&g_stats.n_ellipse, nshape[1]);
    stats_add(&g_stats.n_banana, nshape[2]);
    free(numerator); free(denominator); free(perturbation);
}

/* proc_oa.F90:491-825 for one (level, variable): innovations, multi_scan loop with the
 * surface radius logic (:663-707), innovations after each scan (:748-780), clean_rh (:803). */
int ogo_oa_slab(float* gridded, int iew, int jns, int var, float pressure,
                const float* obs_value, const float* xob, const float* yob, int n,
                float scale, const int* radius_influence, float radius_influence_sfc_mult,
                float dxd, const float* u_banana, const float* v_banana, int passes,
                int smooth_type, int use_first_guess, int scale_cressman_rh_decreases,
                int* n_scans_done)
{
    float* diff = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float* fg_value = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) fg_value[i] = -9999999.0f; /* :514 */
    int rc = OG_OK, scans = 0;

    ogo_innovation(gridded, iew, jns, xob, yob, obs_value, n, scale, use_first_guess, diff,
                   fg_value);

    for (int num_scan = 1; num_scan <= OG_MAX_SCAN; ++num_scan) {
        if (radius_influence[num_scan - 1] < 0) break; /* :643 */
        float sf;
        int is_sfc = fabsf(pressure - 1001.0f) < 0.0001f;
        if (is_sfc) sf = radius_influence_sfc_mult; else sf = 1.00f;
        float r_scaled = (float)radius_influence[num_scan - 1] * sf;
        float r_scaled_scan1 = (float)radius_influence[0] * sf;
        if (radius_influence_sfc_mult < 0.99999f) {
            const float max_cells = 100.0f, min_cells = 4.5f;
            if (is_sfc) {
                if (r_scaled_scan1 > max_cells) r_scaled = r_scaled * (max_cells / r_scaled_scan1);
                if (r_scaled < min_cells) {
                    if (num_scan == 1) { rc = OG_ERR_RADIUS; } /* reference STOPs, :702 */
                    break;
                }
            }
        }
        ogo_cressman(diff, xob, yob, n, gridded, iew, jns, 1, var, f_nint(r_scaled), dxd,
                     u_banana, v_banana, pressure, passes, smooth_type, use_first_guess,
                     fg_value, scale_cressman_rh_decreases);
        ++scans;
        /* more_than_1_scan (max_scan = 10 > 1): always */
        ogo_innovation(gridded, iew, jns, xob, yob, obs_value, n, scale, use_first_guess, diff,
                       NULL);
    }
    if (var == OG_VAR_RH) ogo_clean_rh(gridded, iew, jns, 0.f, 100.f);
    if (n_scans_done) *n_scans_done = scans;
    free(diff); free(fg_value);
    return rc;
}

/* ------------------------------------------------------------------------------- */
/* qc0_module.F90:9-34 */
void ogo_modify_qc_flag(int* qc_flag, int n, const float* err, const float* thr, int test)
{
    for (int i = 0; i < n; ++i) {
        if (fabsf(err[i]) > thr[i]) {
            int qc_small = qc_flag[i] % test;
            int qc_large = (qc_flag[i] / (test * 2)) * 2;
            qc_flag[i] = qc_large * test + qc_small + test;
        }
    }
}

/* qc0_module.F90:40-76 */
void ogo_add_to_qc_flag(int* qc_flag, int to_add)
{
    int qc_small = *qc_flag % to_add;
    int qc_large = (*qc_flag / (to_add * 2)) * 2;
    *qc_flag = qc_large * to_add + qc_small + to_add;
}

/* obs_sort_module.F90:4515-4560 */
static int contains_2n(int sumtocheck, int numtofind)
{
    int curn = 18;
    int sumnow = sumtocheck;
    while (sumnow > 0) {
        if (sumnow < numtofind) return 0;
        else if (sumnow == numtofind) return 1;
        int cur2n = 1 << curn;
        while (cur2n > sumnow) { curn = curn - 1; cur2n = 1 << curn; }
        if (cur2n == numtofind) return 1;
        sumnow = sumnow - cur2n;
    }
    return 0;
}
int ogo_contains_2n(int s, int f) { return contains_2n(s, f); }

/* qc0_module.F90:236-267 */
void ogo_local_time(int time_hhmm, const float* lonob, int n, int* local_time)
{
    for (int i = 0; i < n; ++i) {
        int lt = (int)((float)(time_hhmm / 100) + lonob[i] / 15.f);
        if (lt < 0) lt = 24 + lt;
        local_time[i] = (lt % 24) * 100;
    }
}

/* proc_qc.F90:315-319, ob_access_module.F90:476-486; constants.inc:6-10 */
float ogo_dewpoint(float t, float rh)
{
    const float L_over_Rv = 5418.12f;
    const float Rv_over_L = 1.f / L_over_Rv;
    const float very_small_rh = 0.0001f;
    if (rh < very_small_rh) return 1.f / (1.f / t - Rv_over_L * logf(very_small_rh / 100.f));
    return 1.f / (1.f / t - Rv_over_L * logf(rh / 100.f));
}

/* The LOG of the lines above as the reference build evaluates it: gfortran's LOG(REAL(4)) is a call
 * of the host libm's logf.  Exposed so that the tests can hold the device's restatement of that
 * routine (obsgrid_b200/csrc/og_logf.h) to it over whole argument ranges. */
void ogo_logf_array(const float* x, int n, float* out)
{
    for (int i = 0; i < n; ++i) out[i] = logf(x[i]);
}

/* qc1_module.F90:9-300 */
void ogo_error_max(const float* obs, const float* xob, const float* yob,
                   const float* station_elevation, int* qc_flag, int n,
                   const float* gridded, int iew, int jns, int var, float max_difference,
                   float pressure, const int* local_time,
                   float* aob_out, float* err_out, float* thr_out)
{
    (void)jns;
    size_t nn = (size_t)(n > 0 ? n : 1);
    float* err = (float*)malloc(sizeof(float) * nn);
    float* factor = (float*)malloc(sizeof(float) * nn);
    float* fd = (float*)malloc(sizeof(float) * nn);
    float* aob = (float*)malloc(sizeof(float) * nn);
    float plevel = pressure;

    for (int i = 0; i < n; ++i) factor[i] = 1.0f;
    switch (var) { /* :109-172 */
    case OG_VAR_UU: case OG_VAR_VV:
        if (plevel <= 499.9f) { for (int i = 0; i < n; ++i) factor[i] = 1.5f; }
        else if (plevel <= 1000.1f) { for (int i = 0; i < n; ++i) factor[i] = 1.25f; }
        break;
    case OG_VAR_TT:
        if (plevel > 1000.1f) {
            for (int i = 0; i < n; ++i)
                factor[i] = (local_time[i] >= 1100 && local_time[i] <= 2000) ? 1.5f : 1.25f;
        } else if ((plevel <= 1000.1f) && (plevel > 700.0f)) {
            for (int i = 0; i < n; ++i) factor[i] = 0.85f;
        } else {
            for (int i = 0; i < n; ++i) factor[i] = 0.65f;
        }
        break;
    case OG_VAR_PMSLPSFC:
        for (int i = 0; i < n; ++i) factor[i] = 1.0f + (fabsf(station_elevation[i]) / 2000.0f);
        break;
    case OG_VAR_DEWPOINT:
        for (int i = 0; i < n; ++i) {
            if ((obs[i] < 255.0f) && (obs[i] >= 235.0f)) factor[i] = 1.25f;
            else if ((obs[i] < 235.0f) && (obs[i] >= 215.0f)) factor[i] = 1.50f;
            else if (obs[i] < 215.0f) factor[i] = 1.75f;
            else factor[i] = 1.00f;
        }
        break;
    default: break; /* RH, PMSL */
    }
    for (int i = 0; i < n; ++i) aob[i] = bilinear(gridded, iew, xob[i], yob[i]); /* :176-206 */
    for (int i = 0; i < n; ++i) err[i] = obs[i] - aob[i];                          /* :213 */
    if (var == OG_VAR_UU || var == OG_VAR_VV) {                                     /* :228 */
        for (int i = 0; i < n; ++i) fd[i] = fmaxf(factor[i] * max_difference, 0.15f * aob[i]);
    } else {
        for (int i = 0; i < n; ++i) fd[i] = factor[i] * max_difference;             /* :239 */
    }
    ogo_modify_qc_flag(qc_flag, n, err, fd, OG_FAILS_ERROR_MAX);                    /* :248 */
    if (aob_out) memcpy(aob_out, aob, sizeof(float) * (size_t)n);
    if (err_out) memcpy(err_out, err, sizeof(float) * (size_t)n);
    if (thr_out) memcpy(thr_out, fd, sizeof(float) * (size_t)n);
    free(err); free(factor); free(fd); free(aob);
}

/* ---- binned station loops of buddy_check (fast path, see ogo_set_fast_paths) ---------------
 * station_loop_2 / _3 of qc2_module.F90:223-329 with the pairs restricted to spatial bins one
 * buddy range wide.  Same operations per pair, same order per sum:
 *   pass 1  the as-written loop nest is "for num: for num3 ascending: sum(num) += diff(num3)";
 *           here the nest is swapped -- "for num3 ascending: for every num in the 3x3 bins around
 *           num3" -- so each target still receives its buddies' diffs in ascending num3.
 *   pass 2  the removal loop (:272-286) can only fire for buddies with diff > 1.25*difmax
 *           ("suspects"; difmax either plain or relaxed by :327); per station, in ascending
 *           station order so that the x1.2 relaxation of earlier stations is seen, the suspects of
 *           the 3x3 bins are visited in ascending index. */
typedef struct {
    int numobs; const float* xob; const float* yob; const float* diff_all; float dx, range;
    const int* bin_start; const int* bin_list; int nbx, nby;
    /* targets in bin order (list position e): station index and coordinates, count and sum */
    const float* l_x; const float* l_y; int* l_cnt; float* l_sum;
    int* next_bin;                   /* shared work counter: one target bin per grab */
    int64_t pairs;
    char pad[64];                    /* one cache line per thread's counters */
} buddy_band;

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

/* One candidate num3 against the nt targets of a bin (qc2_module.F90:233-246 for every num of the
 * bin): branch-free so that the compiler can vectorise it; every lane performs the as-written
 * operations (the compiler may not reassociate or contract: no fast-math, -ffp-contract=off). */
__attribute__((optimize("O3", "tree-vectorize", "no-trapping-math", "no-math-errno"), target_clones("avx2", "default")))
static void buddy_targets_of_bin(int nt, const int* restrict t_idx, const float* restrict t_x,
                                 const float* restrict t_y, int* restrict t_cnt,
                                 float* restrict t_sum, int num3, float x3, float y3, float d3,
                                 float dx, float range)
{
    for (int e = 0; e < nt; ++e) {
        float x = t_x[e] - x3;          /* xob(num) - xob(num3) */
        float y = t_y[e] - y3;
        float distance = sqrtf(x * x + y * y) * dx;
        int is_buddy = (distance <= range) & (distance >= 0.1f) & (t_idx[e] != num3);
        t_cnt[e] = t_cnt[e] + is_buddy;
        t_sum[e] = is_buddy ? t_sum[e] + d3 : t_sum[e];
    }
}

static void* buddy_pass1_band(void* arg)
{
    buddy_band* B = (buddy_band*)arg;
    const int nbx = B->nbx, nby = B->nby, nbin = nbx * nby;
    int cap = 4096;
    int* cand = (int*)malloc(sizeof(int) * (size_t)cap);
    int ccap = 0;
    float *c_x = NULL, *c_y = NULL, *c_d = NULL;
    int64_t pairs = 0;
    for (;;) {
        const int c = __atomic_fetch_add(B->next_bin, 1, __ATOMIC_RELAXED);
        if (c >= nbin) break;
        const int t0 = B->bin_start[c], nt = B->bin_start[c + 1] - t0;
        if (nt == 0) continue;
        /* the stations of the 3x3 bins around c, ascending: every target of c meets its buddies in
         * the order of the as-written num3 loop */
        const int bx = c % nbx, by = c / nbx;
        int nc = 0;
        for (int cy = imax(by - 1, 0); cy <= imin(by + 1, nby - 1); ++cy)
            for (int cx = imax(bx - 1, 0); cx <= imin(bx + 1, nbx - 1); ++cx) {
                const int q = cy * nbx + cx, m = B->bin_start[q + 1] - B->bin_start[q];
                if (nc + m > cap) { while (nc + m > cap) cap *= 2; cand = (int*)realloc(cand, sizeof(int) * (size_t)cap); }
                memcpy(cand + nc, B->bin_list + B->bin_start[q], sizeof(int) * (size_t)m);
                nc += m;
            }
        qsort(cand, (size_t)nc, sizeof(int), cmp_int);
        if (nc > ccap) {
            ccap = nc + nc / 4;
            c_x = (float*)realloc(c_x, sizeof(float) * (size_t)ccap);
            c_y = (float*)realloc(c_y, sizeof(float) * (size_t)ccap);
            c_d = (float*)realloc(c_d, sizeof(float) * (size_t)ccap);
        }
        for (int q = 0; q < nc; ++q) { c_x[q] = B->xob[cand[q]]; c_y[q] = B->yob[cand[q]]; c_d[q] = B->diff_all[cand[q]]; }
        /* targets in chunks that stay in the L1 cache while the candidates stream by */
        for (int e0 = 0; e0 < nt; e0 += 1024) {
            const int m = imin(1024, nt - e0), o = t0 + e0;
            for (int q = 0; q < nc; ++q)
                buddy_targets_of_bin(m, B->bin_list + o, B->l_x + o, B->l_y + o, B->l_cnt + o,
                                     B->l_sum + o, cand[q], c_x[q], c_y[q], c_d[q], B->dx, B->range);
        }
        pairs += (int64_t)nc * nt - nt;
    }
    free(cand); free(c_x); free(c_y); free(c_d);
    B->pairs = pairs;
    return NULL;
}


static void buddy_station_loops_binned(const float* obs, int numobs, const float* xob,
                                       const float* yob, int* qc_flag, const float* aob,
                                       float* difmax, int iew, int jns, float dx, float range,
                                       float* err, int* buddy_num, int* buddy_num_final,
                                       int64_t* pairs_examined)
{
    size_t nn = (size_t)(numobs > 0 ? numobs : 1);
    /* |x| <= sqrtf(x*x + y*y) up to an ulp: a pair within range is at most `reach` cells apart in
     * either coordinate, so its two stations sit in the same or in adjacent bins */
    const float reach = range / dx * 1.001f + 1.f;
    const int bs = imax(1, (int)ceilf(reach));
    const int nbx = iew / bs + 2, nby = jns / bs + 2;
    const int nbin = nbx * nby;
    int* bin_of = (int*)malloc(sizeof(int) * nn);
    int* bin_start = (int*)calloc((size_t)nbin + 1, sizeof(int));
    int* bin_list = (int*)malloc(sizeof(int) * nn);
    float* diff_all = (float*)malloc(sizeof(float) * nn);
    float* sum = (float*)calloc(nn, sizeof(float));
    float* l_x = (float*)malloc(sizeof(float) * nn);
    float* l_y = (float*)malloc(sizeof(float) * nn);
    float* l_sum = (float*)calloc(nn, sizeof(float));
    int* l_cnt = (int*)calloc(nn, sizeof(int));
    for (int i = 0; i < numobs; ++i) {
        int bx = imin(imax((int)(xob[i] / (float)bs), 0), nbx - 1);
        int by = imin(imax((int)(yob[i] / (float)bs), 0), nby - 1);
        bin_of[i] = by * nbx + bx;
        bin_start[bin_of[i] + 1]++;
        diff_all[i] = obs[i] - aob[i];          /* :243 for every station as a buddy */
        buddy_num[i] = 0;
    }
    for (int c = 0; c < nbin; ++c) bin_start[c + 1] += bin_start[c];
    int* cursor = (int*)malloc(sizeof(int) * ((size_t)nbin + 1));
    memcpy(cursor, bin_start, sizeof(int) * ((size_t)nbin + 1));
    for (int i = 0; i < numobs; ++i) {                                    /* ascending inside a bin */
        const int e = cursor[bin_of[i]]++;
        bin_list[e] = i; l_x[e] = xob[i]; l_y[e] = yob[i];
    }

    /* pass 1 */
    const int nthr = (g_inner_threads > 1 && numobs >= 4096) ? g_inner_threads : 1;
    buddy_band* bands = (buddy_band*)calloc((size_t)nthr, sizeof(buddy_band));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthr);
    int next_bin = 0;
    for (int t = 0; t < nthr; ++t) {
        bands[t] = (buddy_band){numobs, xob, yob, diff_all, dx, range, bin_start, bin_list, nbx, nby,
                                l_x, l_y, l_cnt, l_sum, &next_bin, 0, {0}};
        if (nthr > 1) pthread_create(&th[t], NULL, buddy_pass1_band, &bands[t]);
        else buddy_pass1_band(&bands[t]);
    }
    for (int t = 0; t < nthr; ++t) {
        if (nthr > 1) pthread_join(th[t], NULL);
        *pairs_examined += bands[t].pairs;
    }
    free(bands); free(th);
    for (int e = 0; e < numobs; ++e) { buddy_num[bin_list[e]] = l_cnt[e]; sum[bin_list[e]] = l_sum[e]; }
    free(l_x); free(l_y); free(l_sum); free(l_cnt);

    /* suspects per bin, ascending */
    int* s_start = (int*)calloc((size_t)nbin + 1, sizeof(int));
    int* s_list = (int*)malloc(sizeof(int) * nn);
    {
        int w = 0;
        for (int c = 0; c < nbin; ++c) {
            s_start[c] = w;
            for (int e = bin_start[c]; e < bin_start[c + 1]; ++e) {
                const int k = bin_list[e];
                if (diff_all[k] > fminf(difmax[k], 1.2f * difmax[k]) * 1.25f) s_list[w++] = k;
            }
        }
        s_start[nbin] = w;
    }

    /* the suspects of the 3x3 bins around every bin, merged ascending (shared by the bin's stations) */
    int* m_start = (int*)calloc((size_t)nbin + 1, sizeof(int));
    for (int c = 0; c < nbin; ++c) {
        const int bx = c % nbx, by = c / nbx;
        int m = 0;
        if (bin_start[c + 1] > bin_start[c])
            for (int cy = imax(by - 1, 0); cy <= imin(by + 1, nby - 1); ++cy)
                for (int cx = imax(bx - 1, 0); cx <= imin(bx + 1, nbx - 1); ++cx)
                    m += s_start[cy * nbx + cx + 1] - s_start[cy * nbx + cx];
        m_start[c + 1] = m_start[c] + m;
    }
    int* m_list = (int*)malloc(sizeof(int) * (size_t)(m_start[nbin] > 0 ? m_start[nbin] : 1));
    for (int c = 0; c < nbin; ++c) {
        if (m_start[c + 1] == m_start[c]) continue;
        const int bx = c % nbx, by = c / nbx;
        int w = m_start[c];
        for (int cy = imax(by - 1, 0); cy <= imin(by + 1, nby - 1); ++cy)
            for (int cx = imax(bx - 1, 0); cx <= imin(bx + 1, nbx - 1); ++cx) {
                const int q = cy * nbx + cx, m = s_start[q + 1] - s_start[q];
                memcpy(m_list + w, s_list + s_start[q], sizeof(int) * (size_t)m);
                w += m;
            }
        qsort(m_list + m_start[c], (size_t)(m_start[c + 1] - m_start[c]), sizeof(int), cmp_int);
    }

    /* pass 2, stations in order (:249-327) */
    float average = 0.f;
    for (int num = 0; num < numobs; ++num) {
        float diff_at_obs = obs[num] - aob[num];
        float s = sum[num];
        if (buddy_num[num] > 0) average = s / (float)buddy_num[num];
        if (buddy_num[num] >= 2) {
            int kob = buddy_num[num];
            const int b = bin_of[num];
            for (int q = m_start[b]; q < m_start[b + 1]; ++q) {
                const int num3 = m_list[q];
                if (num3 == num) continue;
                float x = xob[num] - xob[num3];
                float y = yob[num] - yob[num3];
                float distance = sqrtf(x * x + y * y) * dx;
                ++*pairs_examined;
                if (!(distance <= range && distance >= 0.1f)) continue;
                float diff_check_1 = difmax[num3] * 1.25f;
                float diff_check_2 = difmax[num3];
                float diff_numj = fabsf(diff_all[num3] - average);
                if (diff_all[num3] > diff_check_1 && diff_numj > diff_check_2) {
                    kob = kob - 1;
                    s = s - diff_all[num3];
                }
            }
            buddy_num_final[num] = kob;
            if (kob >= 2) {
                average = s / (float)kob;
                err[num] = diff_at_obs - average;
            } else {
                err[num] = 0.f;
                ogo_add_to_qc_flag(&qc_flag[num], OG_NO_BUDDIES);
            }
        } else {
            err[num] = 0.f;
            ogo_add_to_qc_flag(&qc_flag[num], OG_NO_BUDDIES);
        }
        if (buddy_num_final[num] == 2) difmax[num] = 1.2f * difmax[num]; /* :327 */
    }
    free(m_start); free(m_list);
    free(s_start); free(s_list);
    free(bin_of); free(bin_start); free(bin_list); free(cursor); free(diff_all); free(sum);
}

/* qc2_module.F90:9-388 */
void ogo_buddy_check(const float* obs, int numobs, const float* xob, const float* yob,
                     int* qc_flag, const float* station_elevation, const float* gridded,
                     int iew, int jns, int var, float pressure, float dx, float buddy_weight,
                     float buddy_difference, float* aob_out, float* err_out,
                     float* difmax_out, int* bnf_out)
{
    size_t nn = (size_t)(numobs > 0 ? numobs : 1);
    float* err = (float*)malloc(sizeof(float) * nn);
    float* diff = (float*)malloc(sizeof(float) * nn);
    float* difmax = (float*)malloc(sizeof(float) * nn);
    float* diff_at_obs = (float*)malloc(sizeof(float) * nn);
    float* aob = (float*)malloc(sizeof(float) * nn);
    int* buddy_num = (int*)malloc(sizeof(int) * nn);
    int* buddy_num_final = (int*)calloc(nn, sizeof(int)); /* :116 */
    int* buddy_num_index = (int*)malloc(sizeof(int) * nn);
    float plevel = pressure, range;
    float average = 0.f;
    int64_t pair_tests = 0;

    if (plevel <= 201.0f) range = 650.f;        /* :121-127 */
    else if (plevel <= 1000.1f) range = 300.f;
    else range = 500.f;

    switch (var) { /* :131-191 */
    case OG_VAR_UU: case OG_VAR_VV: {
        float v;
        if (plevel < 401.f) v = buddy_difference * 1.7f;
        else if (plevel < 701.f) v = buddy_difference * 1.4f;
        else if (plevel < 851.f) v = buddy_difference * 1.1f;
        else v = buddy_difference;
        for (int i = 0; i < numobs; ++i) difmax[i] = v;
        break; }
    case OG_VAR_PMSLPSFC:
        for (int i = 0; i < numobs; ++i)
            difmax[i] = buddy_difference * (1.0f + (station_elevation[i] / 2000.0f));
        break;
    case OG_VAR_PMSL:
        for (int i = 0; i < numobs; ++i) difmax[i] = buddy_difference;
        break;
    case OG_VAR_TT: {
        float v;
        if (plevel < 401.f) v = buddy_difference * 0.6f;
        else if (plevel < 701.f) v = buddy_difference * 0.8f;
        else v = buddy_difference;
        for (int i = 0; i < numobs; ++i) difmax[i] = v;
        break; }
    case OG_VAR_RH:
        for (int i = 0; i < numobs; ++i) difmax[i] = buddy_difference;
        break;
    case OG_VAR_DEWPOINT:
        for (int i = 0; i < numobs; ++i) {
            float f;
            if ((obs[i] < 255.0f) && (obs[i] >= 235.0f)) f = 1.25f;
            else if ((obs[i] < 235.0f) && (obs[i] >= 215.0f)) f = 1.50f;
            else if (obs[i] < 215.0f) f = 1.75f;
            else f = 1.00f;
            difmax[i] = buddy_difference * f;
        }
        break;
    default:
        for (int i = 0; i < numobs; ++i) difmax[i] = buddy_difference;
        break;
    }
    for (int i = 0; i < numobs; ++i) difmax[i] = difmax[i] * buddy_weight; /* :193 */
    for (int i = 0; i < numobs; ++i) aob[i] = bilinear(gridded, iew, xob[i], yob[i]);

    if (g_buddy_binned) {
        int64_t examined = 0;
        buddy_station_loops_binned(obs, numobs, xob, yob, qc_flag, aob, difmax, iew, jns, dx,
                                   range, err, buddy_num, buddy_num_final, &examined);
        stats_add(&g_stats.buddy_pairs_examined, examined);
        pair_tests = (int64_t)numobs * (numobs > 0 ? numobs - 1 : 0);   /* what the N^2 loop runs */
    } else
    for (int num = 0; num < numobs; ++num) { /* station_loop_2, :223-329 */
        buddy_num[num] = 0;
        for (int num3 = 0; num3 < numobs; ++num3) {
            if (num3 != num) {
                float x = xob[num] - xob[num3];
                float y = yob[num] - yob[num3];
                float distance = sqrtf(x * x + y * y) * dx;
                ++pair_tests;
                if (distance <= range && distance >= 0.1f) {
                    buddy_num[num] = buddy_num[num] + 1;
                    diff[buddy_num[num] - 1] = obs[num3] - aob[num3];
                    buddy_num_index[buddy_num[num] - 1] = num3;
                }
            }
        }
        diff_at_obs[num] = obs[num] - aob[num];
        float sum = 0.f;
        for (int numj = 0; numj < buddy_num[num]; ++numj) sum = sum + diff[numj];
        if (buddy_num[num] > 0) average = sum / (float)buddy_num[num];

        if (buddy_num[num] >= 2) {
            int kob = buddy_num[num];
            for (int numj = 0; numj < buddy_num[num]; ++numj) {
                float diff_check_1 = difmax[buddy_num_index[numj]] * 1.25f;
                float diff_check_2 = difmax[buddy_num_index[numj]];
                float diff_numj = fabsf(diff[numj] - average);
                if (diff[numj] > diff_check_1 && diff_numj > diff_check_2) {
                    kob = kob - 1;
                    sum = sum - diff[numj];
                }
            }
            buddy_num_final[num] = kob;
            if (kob >= 2) {
                average = sum / (float)kob;
                err[num] = diff_at_obs[num] - average;
            } else {
                err[num] = 0.f;
                ogo_add_to_qc_flag(&qc_flag[num], OG_NO_BUDDIES);
            }
        } else {
            err[num] = 0.f;
            ogo_add_to_qc_flag(&qc_flag[num], OG_NO_BUDDIES);
        }
        if (buddy_num_final[num] == 2) difmax[num] = 1.2f * difmax[num]; /* :327 */
    }
    ogo_modify_qc_flag(qc_flag, numobs, err, difmax, OG_FAILS_BUDDY_CHECK); /* :336 */

    if (aob_out) memcpy(aob_out, aob, sizeof(float) * (size_t)numobs);
    if (err_out) memcpy(err_out, err, sizeof(float) * (size_t)numobs);
    if (difmax_out) memcpy(difmax_out, difmax, sizeof(float) * (size_t)numobs);
    if (bnf_out) memcpy(bnf_out, buddy_num_final, sizeof(int) * (size_t)numobs);
    stats_add(&g_stats.buddy_pair_tests, pair_tests);
    free(err); free(diff); free(difmax); free(diff_at_obs); free(aob);
    free(buddy_num); free(buddy_num_final); free(buddy_num_index);
}

/* qc0_module.F90:81-119 */
void ogo_ob_density(const float* xob, const float* yob, float grid_dist, int numobs,
                    float* tobbox, int iew, int jns)
{
#define TB(j, i) tobbox[(size_t)((i)-1) * jns + (size_t)((j)-1)]
    for (int num = 0; num < numobs; ++num) {
        int iob = (int)xob[num], job = (int)yob[num];
        int rr = f_nint(250.f / grid_dist);
        int jobs = imax(job - rr - 1, 1), jobe = imin(job + rr + 1, jns - 1);
        int iobs = imax(iob - rr - 1, 1), iobe = imin(iob + rr + 1, iew - 1);
        for (int j = jobs; j <= jobe; ++j)
            for (int i = iobs; i <= iobe; ++i) {
                float dist = sqrtf(((float)i - xob[num]) * ((float)i - xob[num]) +
                                   ((float)j - yob[num]) * ((float)j - yob[num]));
                if (dist < 250.f / grid_dist) TB(j, i) = TB(j, i) + 1.f;
            }
    }
#undef TB
}

/* qc0_module.F90:123-232 */
void ogo_obs_distance(const float* tobbox, float* distance, int iew, int jns, float dx)
{
#define IX(j, i) ((size_t)((i)-1) * jns + (size_t)((j)-1))
    float domain_diagonal_m = sqrtf((float)(iew - 1) * dx * (float)(iew - 1) * dx +
                                    (float)(jns - 1) * dx * (float)(jns - 1) * dx);
    float domain_diagonal_grid_cells = domain_diagonal_m / dx;
    for (int i = 1; i <= iew; ++i)
        for (int j = 1; j <= jns; ++j) {
            if (tobbox[IX(j, i)] > 0.5f) { distance[IX(j, i)] = 0.0f; continue; }
            float dijmin = domain_diagonal_grid_cells;
            for (int ii = 1; ii <= iew; ++ii)
                for (int jj = 1; jj <= jns; ++jj)
                    if (tobbox[IX(jj, ii)] > 0.5f) {
                        int di = abs(ii - i), dj = abs(jj - j);
                        float dij = sqrtf((float)dj * (float)dj + (float)di * (float)di); /* dij_lookup */
                        if (dij < dijmin) dijmin = dij;
                    }
            distance[IX(j, i)] = dijmin * dx;
        }
#undef IX
}

/* get_fg_at_point_module.F90:36-114 (after latlon_to_ij) */
float ogo_fg_t_at_point(const float* t3d, const float* h3d, int jns, int iew, int kbu,
                        float x_location, float y_location, float height_msl)
{
#define F3(a, j, i, k) ((a)[((size_t)((k)-1) * iew + (size_t)((i)-1)) * jns + (size_t)((j)-1)])
    const float dTdz = -0.0065f;
    if ((x_location < 1.f) || (x_location > (float)(iew - 1)) || (y_location < 1.f) ||
        (y_location > (float)(jns - 1)))
        return OG_MISSING_R;
    int nx[4], ny[4];
    nx[0] = (int)floorf(x_location); ny[0] = (int)floorf(y_location);
    nx[1] = (int)floorf(x_location); ny[1] = (int)ceilf(y_location);
    nx[2] = (int)ceilf(x_location);  ny[2] = (int)floorf(y_location);
    nx[3] = (int)ceilf(x_location);  ny[3] = (int)ceilf(y_location);
    float tnear[4];
    for (int c = 0; c < 4; ++c) {
        int found = 0;
        float h_below = 0, h_above = 0, t_below = 0, t_above = 0;
        for (int k = 2; k <= kbu - 1; ++k) {
            if ((height_msl >= F3(h3d, ny[c], nx[c], k)) && (height_msl <= F3(h3d, ny[c], nx[c], k + 1))) {
                found = 1;
                h_below = F3(h3d, ny[c], nx[c], k); h_above = F3(h3d, ny[c], nx[c], k + 1);
                t_below = F3(t3d, ny[c], nx[c], k); t_above = F3(t3d, ny[c], nx[c], k + 1);
                break;
            }
        }
        if (found) {
            float w = (h_above - height_msl) / (h_above - h_below);
            tnear[c] = w * t_below + (1.f - w) * t_above;
        } else if (height_msl < F3(h3d, ny[c], nx[c], 2)) {
            tnear[c] = F3(t3d, ny[c], nx[c], 2) - (dTdz * (F3(h3d, ny[c], nx[c], 2) - height_msl));
        } else if (height_msl > F3(h3d, ny[c], nx[c], kbu)) {
            return OG_MISSING_R;
        } else {
            return OG_MISSING_R; /* reference STOPs: "should never be entered" */
        }
    }
    float xn = x_location - floorf(x_location);
    float yn = y_location - floorf(y_location);
    return (1.f - xn) * ((1.f - yn) * tnear[0] + yn * tnear[1]) +
           xn * ((1.f - yn) * tnear[2] + yn * tnear[3]);
#undef F3
}

/* ------------------------------------------------------------------------------- */
/* Per-time harness: the level x variable loops fed from the flat observation table.   */

static int not_missing(float r) { return fabsf(r - OG_MISSING_R) > 1.f; } /* obs_sort_module.F90:1017 */

static int in_domain(float x, float y, int iew, int jns) /* proc_ob_access.F90:231-234 */
{
    return (x >= 2.f) && (x <= (float)(iew - 1)) && (y >= 2.f) && (y <= (float)(jns - 1));
}

/* yx2xy of one level: (jns,iew,k) y-fastest -> (iew,jns) x-fastest (yx2xy.F90:19-23) */
static void slab_get(const float* f3, int iew, int jns, int k, float* xy)
{
    const float* yx = f3 + (size_t)k * iew * jns;
    for (int i = 0; i < iew; ++i)
        for (int j = 0; j < jns; ++j) xy[(size_t)j * iew + i] = yx[(size_t)i * jns + j];
}
static void slab_put(float* f3, int iew, int jns, int k, const float* xy)
{
    float* yx = f3 + (size_t)k * iew * jns;
    for (int i = 0; i < iew; ++i)
        for (int j = 0; j < jns; ++j) yx[(size_t)i * jns + j] = xy[(size_t)j * iew + i];
}

typedef struct {
    const og_time_desc* d; const og_qc_params* p; int k;
} qc_level_job;

/* proc_qc.F90:247-406 for one level kp (0-based k). */
static void qc_level_density(const og_time_desc* d, const og_qc_params* p, int k, int job_mask,
                             float* tobbox)
{
    const int iew = d->iew, jns = d->jns, nrep = d->nrep;
    const size_t npts = (size_t)iew * jns;
    if (!p->qc_test_error_max && !p->qc_test_buddy) return;      /* any_checks_to_do, proc_qc.F90:237 */
    float* gridded = (float*)malloc(sizeof(float) * npts);
    float* gt = NULL;
    size_t nn = (size_t)(nrep > 0 ? nrep : 1);
    float* val = (float*)malloc(sizeof(float) * nn);
    float* x = (float*)malloc(sizeof(float) * nn);
    float* y = (float*)malloc(sizeof(float) * nn);
    float* lon = (float*)malloc(sizeof(float) * nn);
    float* elev = (float*)malloc(sizeof(float) * nn);
    int* qc = (int*)malloc(sizeof(int) * nn);
    int* idx = (int*)malloc(sizeof(int) * nn);
    int* lt = (int*)malloc(sizeof(int) * nn);
    const float pk = d->pressure[k];
    const size_t off = (size_t)k * nrep;

    for (int ivar = 1; ivar <= 7; ++ivar) {
        if (k > 0 && (ivar == 5 || ivar == 6)) continue;
        if (ivar == 6) continue;                    /* qc_psfc = .FALSE. (proc_namelist.F90:73) */
        if (ivar == 7 && !p->qc_dewpoint) continue;
        if (p->var_mask && !(p->var_mask & (1 << (ivar - 1)))) continue;
        if (job_mask && !(job_mask & (1 << (ivar - 1)))) continue;
        float error_difference = 0.f, buddy_difference = 0.f;
        const float* obv = NULL; int* qcv = NULL;
        switch (ivar) {
        case 1: error_difference = p->max_error_uv; buddy_difference = p->max_buddy_uv;
                slab_get(d->u, iew, jns, k, gridded); obv = d->ob_u + off; qcv = d->qc_u + off; break;
        case 2: error_difference = p->max_error_uv; buddy_difference = p->max_buddy_uv;
                slab_get(d->v, iew, jns, k, gridded); obv = d->ob_v + off; qcv = d->qc_v + off; break;
        case 3: error_difference = p->max_error_t; buddy_difference = p->max_buddy_t;
                slab_get(d->t, iew, jns, k, gridded); obv = d->ob_t + off; qcv = d->qc_t + off; break;
        case 4:
            if (pk == 300.f) error_difference = p->max_error_rh * 2.f;      /* proc_qc.F90:282-288 */
            else if (pk < 300.f) error_difference = p->max_error_rh * 3.f;
            else error_difference = p->max_error_rh;
            buddy_difference = p->max_buddy_rh;
            slab_get(d->rh, iew, jns, k, gridded); obv = d->ob_rh + off; qcv = d->qc_rh + off; break;
        case 5: error_difference = p->max_error_p; buddy_difference = p->max_buddy_p;
                slab_get(d->slp, iew, jns, 0, gridded);
                for (size_t q = 0; q < npts; ++q) gridded[q] = gridded[q] * 100.f;   /* :296 */
                obv = d->ob_slp; qcv = d->qc_slp; break;
        case 7: error_difference = p->max_error_dewpoint; buddy_difference = p->max_buddy_dewpoint;
                if (!gt) gt = (float*)malloc(sizeof(float) * npts);
                slab_get(d->t, iew, jns, k, gt);
                slab_get(d->rh, iew, jns, k, gridded);
                for (size_t q = 0; q < npts; ++q) gridded[q] = ogo_dewpoint(gt[q], gridded[q]);
                obv = d->ob_rh + off; qcv = d->qc_rh + off; break;
        }
        /* 'get' (proc_ob_access.F90:146-244, request_qc_max = 2**20) */
        int n = 0;
        for (int r = 0; r < nrep; ++r) {
            float value; int q;
            if (d->bogus && d->bogus[r]) continue;          /* proc_ob_access.F90:179-181 */
            if (ivar == 7) {
                float rhv = d->ob_rh[off + r], tv = d->ob_t[off + r];
                if (!(not_missing(rhv) && d->qc_rh[off + r] < (1 << 20) && not_missing(tv) &&
                      d->qc_t[off + r] < (1 << 20))) continue;
                value = ogo_dewpoint(tv, rhv);
                q = d->qc_rh[off + r];
            } else {
                if (!(not_missing(obv[r]) && qcv[r] < (1 << 20))) continue;
                value = obv[r]; q = qcv[r];
            }
            if (!in_domain(d->xob[r], d->yob[r], iew, jns)) continue;
            val[n] = value; x[n] = d->xob[r]; y[n] = d->yob[r]; lon[n] = d->lon[r];
            elev[n] = d->elev[r]; qc[n] = q; idx[n] = r; ++n;
        }
        ogo_local_time(p->time_hhmmss / 100, lon, n, lt);  /* proc_qc.F90:344 */
        if (tobbox && k == 0) ogo_ob_density(x, y, d->dx_km, n, tobbox, iew, jns);   /* :349-351 */
        if (p->qc_test_error_max)
            ogo_error_max(val, x, y, elev, qc, n, gridded, iew, jns, ivar, error_difference, pk,
                          lt, NULL, NULL, NULL);
        if (p->qc_test_buddy)
            ogo_buddy_check(val, n, x, y, qc, elev, gridded, iew, jns, ivar, pk, d->dx_km,
                            p->buddy_weight, buddy_difference, NULL, NULL, NULL, NULL);
        /* 'put' (proc_ob_access.F90:381-403; DEWPOINT stores into rh%qc, ob_access:771-774) */
        for (int m = 0; m < n; ++m) qcv[idx[m]] = qc[m];
        if (ivar == 7 && d->ob_td && d->qc_td)              /* store_ob 'DEWPOINT', ob_access_module.F90:645-657 */
            for (int m = 0; m < n; ++m) { d->ob_td[off + idx[m]] = val[m]; d->qc_td[off + idx[m]] = qc[m]; }
    }
    free(gridded); free(gt); free(val); free(x); free(y); free(lon); free(elev);
    free(qc); free(idx); free(lt);
}

/* qc0_module.F90:313-434 on the table rows.  The dew-point copies of the flags and the td > t
 * rule (:405-424) need the optional dew-point rows (og_time_desc.ob_td / qc_td). */
static void qc_consistency(const og_time_desc* d, int k_begin, int k_end)
{
    const int dew = d->ob_td && d->qc_td;
    for (int k = k_begin; k < k_end; ++k)
        for (int r = 0; r < d->nrep; ++r) {
            size_t q = (size_t)k * d->nrep + r;
            int qc_u = d->qc_u[q], qc_v = d->qc_v[q];
            if (qc_u > qc_v) d->qc_v[q] = qc_u;
            if (qc_v > qc_u) d->qc_u[q] = qc_v;
            int qc_t = d->qc_t[q], qc_rh = d->qc_rh[q];
            int scratch = 0;
            int* qtd = dew ? &d->qc_td[q] : &scratch;
            if (contains_2n(qc_t, OG_FAILS_ERROR_MAX) && !contains_2n(qc_rh, OG_FAILS_ERROR_MAX)) {
                ogo_add_to_qc_flag(&d->qc_rh[q], OG_FAILS_ERROR_MAX);
                ogo_add_to_qc_flag(qtd, OG_FAILS_ERROR_MAX);
            }
            if (contains_2n(qc_t, OG_FAILS_BUDDY_CHECK) && !contains_2n(qc_rh, OG_FAILS_BUDDY_CHECK)) {
                ogo_add_to_qc_flag(&d->qc_rh[q], OG_FAILS_BUDDY_CHECK);
                ogo_add_to_qc_flag(qtd, OG_FAILS_BUDDY_CHECK);
            }
            if (dew && (d->ob_td[q] > d->ob_t[q]) && (qc_rh < OG_FAILS_BUDDY_CHECK)) {   /* :408-422 */
                ogo_add_to_qc_flag(&d->qc_rh[q], OG_FAILS_BUDDY_CHECK);
                ogo_add_to_qc_flag(qtd, OG_FAILS_BUDDY_CHECK);
            }
        }
}

static void qc_level(const og_time_desc* d, const og_qc_params* p, int k, int job_mask)
{
    qc_level_density(d, p, k, job_mask, NULL);
}

/* proc_qc.F90:243-420 for a surface-FDDA time: the level x variable loop with the observation
 * density accumulated at kp = 1, then :416-420. */
int ogo_qc_time_fdda(const og_time_desc* d, const og_qc_params* p, float* tobbox, float* odis)
{
    for (int k = 0; k < d->kbu; ++k) qc_level_density(d, p, k, 0, tobbox);
    const size_t npts = (size_t)d->iew * d->jns;
    for (size_t q = 0; q < npts; ++q) {
        tobbox[q] = tobbox[q] / 5.f;
        if (tobbox[q] < 1.f) tobbox[q] = 0.f;
    }
    ogo_obs_distance(tobbox, odis, d->iew, d->jns, d->dx_km);
    qc_consistency(d, 0, d->kbu);
    return OG_OK;
}

void ogo_qc_consistency(const og_time_desc* d, int k_begin, int k_end)
{
    qc_consistency(d, k_begin, k_end);
}

int ogo_qc_time(const og_time_desc* d, const og_qc_params* p, int k_begin, int k_end,
                int run_consistency)
{
    for (int k = k_begin; k < k_end; ++k) qc_level(d, p, k, 0);
    if (run_consistency) qc_consistency(d, k_begin, k_end);
    return OG_OK;
}

/* proc_oa.F90:390-847 for one (level, variable) slab.  u_banana/v_banana are the
 * pre-analysis winds of the level in (iew,jns) order (proc_oa.F90:376-386). */
static int oa_slab_job(const og_time_desc* d, const og_oa_params* p, int k, int ivar,
                       const float* u_banana, const float* v_banana)
{
    const int iew = d->iew, jns = d->jns, nrep = d->nrep;
    const size_t npts = (size_t)iew * jns;
    if (ivar == 5 && k > 0) return OG_OK;
    if (p->var_mask && !(p->var_mask & (1 << (ivar - 1)))) return OG_OK;
    float* dum2d = (float*)malloc(sizeof(float) * npts);
    size_t nn = (size_t)(nrep > 0 ? nrep : 1);
    float* val = (float*)malloc(sizeof(float) * nn);
    float* x = (float*)malloc(sizeof(float) * nn);
    float* y = (float*)malloc(sizeof(float) * nn);
    const float pk = d->pressure[k];
    const size_t off = (size_t)k * nrep;
    const int qc_max = OG_FAILS_ERROR_MAX < OG_FAILS_BUDDY_CHECK ? OG_FAILS_ERROR_MAX
                                                                 : OG_FAILS_BUDDY_CHECK;
    int passes; float scale = (ivar == 5) ? 100.f : 1.f;
    const int psfc[5] = {p->smooth_sfc_wind, p->smooth_sfc_wind, p->smooth_sfc_temp,
                         p->smooth_sfc_rh, p->smooth_sfc_slp};
    const int pup[5] = {p->smooth_upper_wind, p->smooth_upper_wind, p->smooth_upper_temp,
                        p->smooth_upper_rh, 0};
    passes = (k == 0) ? psfc[ivar - 1] : pup[ivar - 1];
    float* f3 = NULL; const float* obv = NULL; const int* qcv = NULL;
    switch (ivar) {
    case 1: f3 = d->u; obv = d->ob_u + off; qcv = d->qc_u + off; break;
    case 2: f3 = d->v; obv = d->ob_v + off; qcv = d->qc_v + off; break;
    case 3: f3 = d->t; obv = d->ob_t + off; qcv = d->qc_t + off; break;
    case 4: f3 = d->rh; obv = d->ob_rh + off; qcv = d->qc_rh + off; break;
    default: f3 = d->slp; obv = d->ob_slp; qcv = d->qc_slp; break;
    }
    int n = 0; /* 'use' (proc_ob_access.F90:270-355) */
    for (int r = 0; r < nrep; ++r) {
        if (!(not_missing(obv[r]) && qcv[r] < qc_max)) continue;
        if (!in_domain(d->xob[r], d->yob[r], iew, jns)) continue;
        val[n] = obv[r]; x[n] = d->xob[r]; y[n] = d->yob[r]; ++n;
    }
    slab_get(f3, iew, jns, ivar == 5 ? 0 : k, dum2d);
    int rc = ogo_oa_slab(dum2d, iew, jns, ivar, pk, val, x, y, n, scale, p->radius_influence,
                         p->radius_influence_sfc_mult, d->dx_km, u_banana, v_banana, passes,
                         p->smooth_type, p->use_first_guess, p->scale_cressman_rh_decreases,
                         NULL);
    slab_put(f3, iew, jns, ivar == 5 ? 0 : k, dum2d);
    free(dum2d); free(val); free(x); free(y);
    return rc;
}

/* proc_oa.F90:356-849 for one level: variables in the reference's order. */
static int oa_level(const og_time_desc* d, const og_oa_params* p, int k)
{
    const size_t npts = (size_t)d->iew * d->jns;
    float* u_banana = (float*)malloc(sizeof(float) * npts);
    float* v_banana = (float*)malloc(sizeof(float) * npts);
    int rc = OG_OK;
    slab_get(d->u, d->iew, d->jns, k, u_banana); /* :376-386 (pre-analysis winds) */
    slab_get(d->v, d->iew, d->jns, k, v_banana);
    for (int ivar = 1; ivar <= 5; ++ivar) { /* PSFC (6) needs oa_psfc: off */
        int r1 = oa_slab_job(d, p, k, ivar, u_banana, v_banana);
        if (r1 != OG_OK) rc = r1;
    }
    free(u_banana); free(v_banana);
    return rc;
}

/* mixing_ratio's side effect on the first guess (final_analysis_module.F90:1055) */
static void clamp_fg_rh(const og_time_desc* d, int k_begin, int k_end)
{
    for (int k = k_begin; k < k_end; ++k)
        for (int i = 0; i < d->iew - 1; ++i)
            for (int j = 0; j < d->jns - 1; ++j) {
                float* v = &d->rh[((size_t)k * d->iew + i) * d->jns + j];
                *v = fminf(fmaxf(*v, 1.f), 100.f);
            }
}

int ogo_oa_time(const og_time_desc* d, const og_oa_params* p, int k_begin, int k_end)
{
    int rc = OG_OK;
    if (p->clamp_fg_rh) clamp_fg_rh(d, k_begin, k_end);
    for (int k = k_begin; k < k_end; ++k) {
        if (p->surface_only && k > 0) break; /* proc_oa.F90:365 */
        int r = oa_level(d, p, k);
        if (r != OG_OK) rc = r;
    }
    return rc;
}

/* ---- pthread fan-out -------------------------------------------------------------------
 * OA: one job per (level, variable) slab -- slabs are independent once the pre-analysis winds
 * of every level have been snapshotted; QC: one job per level and variable group
 * (DEWPOINT reads the RH flags of its own level).  Jobs over many observations ("big": the
 * surface slabs of configs 3-5) run one after the other, each on all threads through the
 * inner fast paths (ogo_set_fast_paths); the many small ones share a pool, one thread each. */
typedef struct {
    const og_time_desc* d; const og_oa_params* po; const og_qc_params* pq;
    int k_begin, k_end; int* next; int rc;
    const float* ub_all; const float* vb_all;   /* [k_end-k_begin][jns][iew] */
    const unsigned char* big;                   /* [nlev*5]: handled outside the pool */
} mt_job;

static const int k_qc_masks[5] = {1 << 0, 1 << 1, 1 << 2, 1 << 4, (1 << 3) | (1 << 6)};

static int level_ob_count(const og_time_desc* d, int k)
{
    int n = 0;
    const float* t = d->ob_t + (size_t)k * d->nrep;
    for (int r = 0; r < d->nrep; ++r) n += not_missing(t[r]);
    return n;
}

static int oa_job(const mt_job* j, int q)
{
    const size_t npts = (size_t)j->d->iew * j->d->jns;
    int k = j->k_begin + q / 5, ivar = 1 + q % 5;   /* surface (largest) slabs first */
    if (j->po->surface_only && k > 0) return OG_OK;
    return oa_slab_job(j->d, j->po, k, ivar, j->ub_all + npts * (size_t)(k - j->k_begin),
                       j->vb_all + npts * (size_t)(k - j->k_begin));
}

static void* oa_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    const int nlev = j->k_end - j->k_begin;
    for (;;) {
        int q = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (q >= nlev * 5) break;
        if (j->big[q]) continue;
        int r = oa_job(j, q);
        if (r != OG_OK) j->rc = r;
    }
    return NULL;
}
static void* qc_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    const int nlev = j->k_end - j->k_begin;
    for (;;) {
        int q = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (q >= nlev * 5) break;
        if (j->big[q]) continue;
        qc_level(j->d, j->pq, j->k_begin + q / 5, k_qc_masks[q % 5]);
    }
    return NULL;
}

/* big[q] for the jobs of the levels [k_begin,k_end): more than `limit` observations */
static unsigned char* mark_big_jobs(const og_time_desc* d, int k_begin, int k_end, int limit)
{
    const int nlev = k_end - k_begin;
    unsigned char* big = (unsigned char*)calloc((size_t)(nlev > 0 ? nlev : 1) * 5, 1);
    for (int k = k_begin; k < k_end; ++k)
        if (level_ob_count(d, k) > limit)
            for (int v = 0; v < 5; ++v) big[(k - k_begin) * 5 + v] = 1;
    return big;
}

int ogo_oa_time_mt(const og_time_desc* d, const og_oa_params* p, int k_begin, int k_end,
                   int threads)
{
    if (threads <= 1) return ogo_oa_time(d, p, k_begin, k_end);
    if (p->clamp_fg_rh) clamp_fg_rh(d, k_begin, k_end);
    const size_t npts = (size_t)d->iew * d->jns;
    const int nlev = k_end - k_begin;
    if (nlev <= 0) return OG_OK;
    float* ub = (float*)malloc(sizeof(float) * npts * (size_t)nlev);
    float* vb = (float*)malloc(sizeof(float) * npts * (size_t)nlev);
    for (int k = k_begin; k < k_end; ++k) {
        slab_get(d->u, d->iew, d->jns, k, ub + npts * (size_t)(k - k_begin));
        slab_get(d->v, d->iew, d->jns, k, vb + npts * (size_t)(k - k_begin));
    }
    int next = 0, rc = OG_OK;
    unsigned char* big = mark_big_jobs(d, k_begin, k_end, 30000);
    const int saved_inner = g_inner_threads;
    mt_job seq = {d, p, NULL, k_begin, k_end, &next, OG_OK, ub, vb, big};
    g_inner_threads = threads;
    for (int q = 0; q < nlev * 5; ++q)
        if (big[q]) { int r = oa_job(&seq, q); if (r != OG_OK) rc = r; }
    g_inner_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (mt_job){d, p, NULL, k_begin, k_end, &next, OG_OK, ub, vb, big};
        pthread_create(&th[t], NULL, oa_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    g_inner_threads = saved_inner;
    free(th); free(jobs); free(ub); free(vb); free(big);
    return rc;
}

int ogo_qc_time_mt(const og_time_desc* d, const og_qc_params* p, int k_begin, int k_end,
                   int run_consistency, int threads)
{
    if (threads <= 1) return ogo_qc_time(d, p, k_begin, k_end, run_consistency);
    int next = 0;
    const int nlev = k_end - k_begin;
    unsigned char* big = mark_big_jobs(d, k_begin, k_end, 30000);
    const int saved_inner = g_inner_threads;
    if (!g_buddy_binned) memset(big, 0, (size_t)(nlev > 0 ? nlev : 1) * 5);   /* the N^2 loops do not split */
    g_inner_threads = threads;
    for (int q = 0; q < nlev * 5; ++q)
        if (big[q]) qc_level(d, p, k_begin + q / 5, k_qc_masks[q % 5]);
    g_inner_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (mt_job){d, NULL, p, k_begin, k_end, &next, OG_OK, NULL, NULL, big};
        pthread_create(&th[t], NULL, qc_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    g_inner_threads = saved_inner;
    free(th); free(jobs); free(big);
    if (run_consistency) qc_consistency(d, k_begin, k_end);
    return OG_OK;
}

/* ======================================================================================
 * Post-analysis epilogue: the per-grid-point maps between proc_oa and the
 * netCDF writer.  3-D arrays are (jns,iew,kbu), j fastest, as first_guess.inc declares them.
 * ====================================================================================== */
#define IX3(j, i, k, jns, iew) (((size_t)(k) * (size_t)(iew) + (size_t)(i)) * (size_t)(jns) + (size_t)(j))

/* final_analysis_module.F90:316-319 with crs2dot_pertU (first_guess_module.F90:59-74):
 * pertubation = u - uA; averaged along dim2 (iew) onto the staggered points; u = uC + pert. */
void ogo_wind_increment_u(float* u, const float* uA, const float* uC, int jns, int iew, int kbu)
{
    const size_t n = (size_t)jns * iew * kbu;
    float* pert = (float*)malloc(sizeof(float) * n);
    float* dummy = (float*)malloc(sizeof(float) * n);
    for (size_t q = 0; q < n; ++q) pert[q] = u[q] - uA[q];
    for (int k = 0; k < kbu; ++k)
        for (int i = 0; i < iew; ++i)
            for (int j = 0; j < jns; ++j) {
                float v;
                if (i == 0) v = pert[IX3(j, 0, k, jns, iew)];                       /* :70 */
                else if (i == iew - 1) v = pert[IX3(j, iew - 2, k, jns, iew)];      /* :71 */
                else v = (pert[IX3(j, i - 1, k, jns, iew)] + pert[IX3(j, i, k, jns, iew)]) * 0.5f; /* :68-69 */
                dummy[IX3(j, i, k, jns, iew)] = v;
            }
    for (size_t q = 0; q < n; ++q) u[q] = uC[q] + dummy[q];
    free(pert); free(dummy);
}

/* final_analysis_module.F90:336-339 with crs2dot_pertV (first_guess_module.F90:76-91): along dim1 (jns). */
void ogo_wind_increment_v(float* v, const float* vA, const float* vC, int jns, int iew, int kbu)
{
    const size_t n = (size_t)jns * iew * kbu;
    float* pert = (float*)malloc(sizeof(float) * n);
    float* dummy = (float*)malloc(sizeof(float) * n);
    for (size_t q = 0; q < n; ++q) pert[q] = v[q] - vA[q];
    for (int k = 0; k < kbu; ++k)
        for (int i = 0; i < iew; ++i)
            for (int j = 0; j < jns; ++j) {
                float w;
                if (j == 0) w = pert[IX3(0, i, k, jns, iew)];                       /* :87 */
                else if (j == jns - 1) w = pert[IX3(jns - 2, i, k, jns, iew)];      /* :88 */
                else w = (pert[IX3(j - 1, i, k, jns, iew)] + pert[IX3(j, i, k, jns, iew)]) * 0.5f; /* :85-86 */
                dummy[IX3(j, i, k, jns, iew)] = w;
            }
    for (size_t q = 0; q < n; ++q) v[q] = vC[q] + dummy[q];
    free(pert); free(dummy);
}

/* final_analysis_module.F90:364-370: surface level of PRES follows the SLP increment. */
void ogo_pres_sfc_from_slp(const float* pres_sfc, const float* slp_C, const float* slp_x,
                           float* out, int jns, int iew)
{
    const size_t n = (size_t)jns * iew;
    for (size_t q = 0; q < n; ++q)
        out[q] = pres_sfc[q] + (pres_sfc[q] / slp_C[q]) * (slp_x[q] - slp_C[q]);
}

/* es of mixing_ratio (:1058): svp1*10. is a PARAMETER product, folded in FP32 */
static float sat_vapor_pressure(float t)
{
    const float svp1x10 = 0.6112f * 10.f, svp2 = 17.67f, svp3 = 29.65f, svpt0 = 273.15f;
    return svp1x10 * expf(svp2 * (t - svpt0) / (t - svp3));
}

/* sfcprs, final_analysis_module.F90:1087-1201.  Returns OG_ERR_INVALID when the 850/700/500 hPa
 * levels are missing (the reference STOPs, :1122-1127). */
static int sfcprs(const float* t, const float* q, const float* h, const float* slp_x,
                  const float* terrain, const float* pressure, int iew, int jns, int kbu,
                  float* psfc)
{
    const float R = 287.f, g = 9.806f, gamma = 6.5e-3f;
    const float gammarg = gamma * R / g;
    int k850 = -1, k700 = -1, k500 = -1;
    for (int k = 1; k < kbu; ++k) {
        const long p = lroundf(pressure[k]);
        if (p == 850) k850 = k; else if (p == 700) k700 = k; else if (p == 500) k500 = k;
    }
    if (k850 < 0 || k700 < 0 || k500 < 0) return OG_ERR_INVALID;
    const float p78 = 1.f / logf(850.f / 700.f), p57 = 1.f / logf(700.f / 500.f);
    for (int i = 0; i < iew - 1; ++i)
        for (int j = 0; j < jns - 1; ++j) {
            const size_t o = (size_t)i * jns + j;
            const float ht = -terrain[o] / h[IX3(j, i, k850, jns, iew)];            /* :1133 */
            const float ps0 = slp_x[o] * powf(slp_x[o] / 85000.f, ht);              /* :1134 */
            float p1;                                                               /* :1142-1148 */
            if (ps0 > 95000.f) p1 = 85000.f;
            else if (ps0 > 70000.f) p1 = ps0 - 10000.f;
            else p1 = 50000.f;
            const float t850 = t[IX3(j, i, k850, jns, iew)] * (1.f + 0.608f * q[IX3(j, i, k850, jns, iew)]);
            const float t700 = t[IX3(j, i, k700, jns, iew)] * (1.f + 0.608f * q[IX3(j, i, k700, jns, iew)]);
            const float t500 = t[IX3(j, i, k500, jns, iew)] * (1.f + 0.608f * q[IX3(j, i, k500, jns, iew)]);
            const float gamma78 = logf(t850 / t700) * p78, gamma57 = logf(t700 / t500) * p57;
            float t1;                                                               /* :1173-1181 */
            if (ps0 > 95000.f) t1 = t850;
            else if (ps0 > 85000.f) t1 = t700 * powf(p1 / 70000.f, gamma78);
            else if (ps0 > 70000.f) t1 = t500 * powf(p1 / 50000.f, gamma57);
            else t1 = t500;
            const float tslv = t1 * powf(slp_x[o] / p1, gammarg);                    /* :1185 */
            const float tsfc = tslv - gamma * terrain[o];                           /* :1189 */
            psfc[o] = slp_x[o] * expf((-terrain[o] * g) / (R / 2.f * (tsfc + tslv))); /* :1197 */
        }
    return OG_OK;
}

/* mixing_ratio, final_analysis_module.F90:1026-1083: clamps rh in place (:1055), fills qv on the
 * cross points (1:jns-1, 1:iew-1) of every level and psfc (Pa) on the same points. */
int ogo_mixing_ratio(float* rh, const float* t, const float* h, const float* terrain,
                     const float* slp_x, const float* pressure, int iew, int jns, int kbu,
                     float* qv, float* psfc)
{
    const float eps = 0.622f;
    for (int k = 0; k < kbu; ++k)
        for (int i = 0; i < iew - 1; ++i)
            for (int j = 0; j < jns - 1; ++j) {
                float* r = &rh[IX3(j, i, k, jns, iew)];
                *r = fminf(fmaxf(*r, 1.f), 100.f);
            }
    for (int k = 1; k < kbu; ++k)
        for (int i = 0; i < iew - 1; ++i)
            for (int j = 0; j < jns - 1; ++j) {
                const size_t o = IX3(j, i, k, jns, iew);
                const float es = sat_vapor_pressure(t[o]);
                const float qs = eps * es / (pressure[k] - es);
                qv[o] = fmaxf(0.01f * rh[o] * qs, 0.0f);
            }
    int rc = sfcprs(t, qv, h, slp_x, terrain, pressure, iew, jns, kbu, psfc);
    if (rc != OG_OK) return rc;
    for (int i = 0; i < iew - 1; ++i)
        for (int j = 0; j < jns - 1; ++j) {
            const size_t o = IX3(j, i, 0, jns, iew);
            const float es = sat_vapor_pressure(t[o]);
            const float qs = eps * es / (psfc[o] / 100.f - es);
            qv[o] = fmaxf(0.01f * rh[o] * qs, 0.0f);
        }
    return OG_OK;
}

/* temporal_interp, tinterp_module.F90:53-215, for one field.  Returns OG_ERR_INVALID where the
 * reference STOPs on a zero weight sum (:112-115). */
int ogo_temporal_interp(const float* first, const float* second, float* out, int jns, int iew,
                        int nlev, int icount_fdda, int icount_1, int icount_2, int cross)
{
    int w1i = icount_2 - icount_fdda, w2i = icount_fdda - icount_1;      /* :98-104 */
    if (w1i == 0 && w2i == 0) w1i = 1;                                    /* :106-109 */
    if (w1i + w2i == 0) return OG_ERR_INVALID;
    const float w1 = (float)w1i, w2 = (float)w2i, wd = (float)(w1i + w2i);
    const int je = cross ? jns - 1 : jns, ie = cross ? iew - 1 : iew;
    for (int k = 0; k < nlev; ++k) {
        for (int i = 0; i < ie; ++i)
            for (int j = 0; j < je; ++j) {
                const size_t o = IX3(j, i, k, jns, iew);
                out[o] = (w1 * first[o] + w2 * second[o]) / wd;          /* :128-130 */
            }
        if (cross) {
            for (int i = 0; i < iew - 1; ++i)                             /* t(jns,:iew-1,:) = t(jns-1,:iew-1,:) */
                out[IX3(jns - 1, i, k, jns, iew)] = out[IX3(jns - 2, i, k, jns, iew)];
            for (int j = 0; j < jns; ++j)                                 /* t(:jns,iew,:) = t(:jns,iew-1,:) */
                out[IX3(j, iew - 1, k, jns, iew)] = out[IX3(j, iew - 2, k, jns, iew)];
        }
    }
    return OG_OK;
}
