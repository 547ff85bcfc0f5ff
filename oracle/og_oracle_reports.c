/*
 * og_oracle_reports.c -- CPU oracle, part 2: the reference's observation access AS WRITTEN, slab by
 * slab over the report list, for checking the flat-table path (include/obsgrid_reports.h).
 *
 * TEST INFRASTRUCTURE ONLY (see og_oracle.h).  Restates, on the plain-array image of the report
 * list (og_report / og_meas: the `surface` linked list flattened in order), the routines the table
 * builder replaces with one pass per analysis time:
 *   query_ob        src/ob_access_module.F90:11-535
 *   store_ob        src/ob_access_module.F90:538-817
 *   proc_ob_access  src/proc_ob_access.F90:146-403 ('get', 'use', 'put')
 *   qc_consistency  src/qc0_module.F90:313-434 (over every measurement of every report)
 * and the level x variable loops of proc_qc (src/proc_qc.F90:243-408) / the 'use' selection of
 * proc_oa (src/proc_oa.F90:455-470) that call them.  latlon_to_ij is the caller's (og_report.x, .y).
 */
#define _GNU_SOURCE
#include "og_oracle.h"
#include "../include/obsgrid_reports.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int not_missing(float r) { return fabsf(r - OG_MISSING_R) > 1.f; }      /* ob_access_module.F90:65 */
static int eps_equal(float a, float b, float eps) { return fabsf(a - b) < eps; }

/* geth_newdate (date_pack_module.F90) as far as query_ob needs it: "CCYY-MM-DD_HH:mm:ss" plus a
 * number of seconds, through a day count. */
static long long civil_days(long long y, int m, int d)
{
    y -= m <= 2;
    long long era = (y >= 0 ? y : y - 399) / 400;
    int yoe = (int)(y - era * 400);
    int doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    int doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}
static void civil_from_days(long long z, int* y, int* m, int* d)
{
    z += 719468;
    long long era = (z >= 0 ? z : z - 146096) / 146097;
    int doe = (int)(z - era * 146097);
    int yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    int mp = (5 * doy + 2) / 153;
    *d = doy - (153 * mp + 2) / 5 + 1;
    *m = mp + (mp < 10 ? 3 : -9);
    *y = (int)(yoe + era * 400 + (*m <= 2));
}
static void newdate(char out[64], int date, int time, int ds)
{
    long long s = civil_days(date / 10000, (date / 100) % 100, date % 100) * 86400LL +
                  (time / 10000) * 3600 + ((time / 100) % 100) * 60 + time % 100 + ds;
    long long day = s >= 0 ? s / 86400 : -((-s + 86399) / 86400);
    int sec = (int)(s - day * 86400), y, m, d;
    civil_from_days(day, &y, &m, &d);
    snprintf(out, 64, "%04d-%02d-%02d_%02d:%02d:%02d", y, m, d, sec / 3600, (sec / 60) % 60, sec % 60);
}

static int window_ds(const char* platform)         /* ob_access_module.F90:88-112 */
{
    int special = !strncmp(platform, "FM-97 AIREP", 11) || !strncmp(platform, "FM-36 TEMP SHIP", 15) ||
                  !strncmp(platform, "FM-35 TEMP", 10) || !strncmp(platform, "FM-88 SATOB", 11);
    if (!strncmp(platform + 35, "    ", 4)) return special ? 3600 : 1800;
    char f[5]; memcpy(f, platform + 35, 4); f[4] = 0;
    /* READ (..., FMT='(I4)'): blanks ignored, optional sign, digits only */
    char g[5]; int n = 0;
    for (int i = 0; i < 4; ++i) if (f[i] != ' ') g[n++] = f[i];
    g[n] = 0;
    int ok = n > 0, i0 = (g[0] == '+' || g[0] == '-') ? 1 : 0;
    if (i0 == n) ok = 0;
    for (int i = i0; i < n; ++i) if (g[i] < '0' || g[i] > '9') ok = 0;
    if (!ok) return special ? 3600 : 1800;
    return atoi(g);
}

static float dewpoint_of(float t_value, float rh_value)     /* ob_access_module.F90:206-210 */
{
    return ogo_dewpoint(t_value, rh_value);
}

/* query_ob.  var: og variable code (5 = PMSL, 7 = DEWPOINT; PSFC / PMSLPSFC are not requested by the
 * loops restated here, qc_psfc = oa_psfc = .FALSE.).  qc stays OG_MISSING_I when nothing is found. */
void ogo_query_ob(og_report* obs, og_meas* meas, int date, int time, int var, float request_level,
                  int request_qc_max, int request_p_diff, float* value, int* qc)
{
    char obs_date[20], p1[64], m1[64];
    const char* c = obs->date_char;
    snprintf(obs_date, sizeof obs_date, "%.4s-%.2s-%.2s_%.2s:%.2s:%.2s", c, c + 4, c + 6, c + 8, c + 10, c + 12);
    int ds = window_ds(obs->platform);
    newdate(p1, date, time, ds);
    newdate(m1, date, time, -ds);
    int close = (strcmp(obs_date, p1) <= 0) && (strcmp(obs_date, m1) >= 0);
    if (!close) {                                              /* :511-518 */
        if (obs->n_meas > 0) {
            og_meas* s = &meas[obs->first_meas];
            s->pressure.qc = s->height.qc = s->temperature.qc = s->u.qc = s->v.qc = s->rh.qc = OG_NO_QC_POSSIBLE;
        }
        return;
    }
    if (var == OG_VAR_PMSL) {                                  /* :131-137 */
        if (not_missing(obs->slp.data) && obs->slp.qc < request_qc_max) { *value = obs->slp.data; *qc = obs->slp.qc; }
        return;
    }
    og_meas* found = NULL;
    if ((int)lroundf(request_level) == 1001) {                 /* search_for_sfc, :155-309 */
        for (int q = 0; q < obs->n_meas; ++q) {
            og_meas* nx = &meas[obs->first_meas + q];
            if (eps_equal(nx->height.data, obs->elevation, 1.f) && !eps_equal(nx->height.data, OG_MISSING_R, 1.f) &&
                !eps_equal(obs->elevation, OG_MISSING_R, 1.f)) { found = nx; break; }
            else if ((nx->height.data - obs->elevation > 10.f) && !eps_equal(nx->height.data, OG_MISSING_R, 1.f) &&
                     !eps_equal(obs->elevation, OG_MISSING_R, 1.f)) break;
        }
    } else {                                                   /* :311-489 */
        int curlev = 0, uselev = -1, maxlev;
        float minpdiff = 9999999.f;
        for (int q = 0; q < obs->n_meas; ++q) {
            curlev = curlev + 1;
            float curpdiff = fabsf(meas[obs->first_meas + q].pressure.data - request_level * 100.f);
            if (curpdiff < (float)request_p_diff) {
                if (curpdiff < minpdiff) { minpdiff = curpdiff; uselev = curlev; }
            }
        }
        maxlev = curlev;
        curlev = 0;
        for (int q = 0; q < obs->n_meas; ++q) {
            og_meas* nx = &meas[obs->first_meas + q];
            curlev = curlev + 1;
            float curpdiff = fabsf(nx->pressure.data - request_level * 100.f);
            if ((((maxlev == 1) && (curlev == uselev)) || ((maxlev != 1) && (curpdiff < 1.f))) &&
                (!eps_equal(nx->height.data, obs->elevation, 1.f) || eps_equal(nx->height.data, OG_MISSING_R, 1.f))) {
                if ((maxlev == 1) && (curlev == uselev)) {
                    if (!ogo_contains_2n(nx->pressure.qc, OG_EXTEND_INFLUENCE))
                        nx->pressure.qc = nx->pressure.qc + OG_EXTEND_INFLUENCE;
                }
                found = nx;
                break;
            }
        }
    }
    if (!found) return;
    switch (var) {                                             /* which_sfc_var / which_var */
    case OG_VAR_UU: if (not_missing(found->u.data) && found->u.qc < request_qc_max) { *value = found->u.data; *qc = found->u.qc; } break;
    case OG_VAR_VV: if (not_missing(found->v.data) && found->v.qc < request_qc_max) { *value = found->v.data; *qc = found->v.qc; } break;
    case OG_VAR_TT: if (not_missing(found->temperature.data) && found->temperature.qc < request_qc_max) { *value = found->temperature.data; *qc = found->temperature.qc; } break;
    case OG_VAR_RH: if (not_missing(found->rh.data) && found->rh.qc < request_qc_max) { *value = found->rh.data; *qc = found->rh.qc; } break;
    case OG_VAR_DEWPOINT:
        if ((not_missing(found->rh.data) && found->rh.qc < request_qc_max) &&
            (not_missing(found->temperature.data) && found->temperature.qc < request_qc_max)) {
            *value = dewpoint_of(found->temperature.data, found->rh.data);
            *qc = found->rh.qc;
        }
        break;
    default: break;
    }
}

/* store_ob; returns 0, or -1 where the reference stops with a fatal error (level not found). */
int ogo_store_ob(og_report* obs, og_meas* meas, int var, float request_level, int request_p_diff, float value, int qc)
{
    if (var == OG_VAR_PMSL) { obs->slp.data = value; obs->slp.qc = qc; return 0; }
    og_meas* found = NULL;
    if ((int)lroundf(request_level) == 1001) {
        for (int q = 0; q < obs->n_meas; ++q) {
            og_meas* nx = &meas[obs->first_meas + q];
            if (eps_equal(nx->height.data, obs->elevation, 1.f) && !eps_equal(nx->height.data, OG_MISSING_R, 1.f) &&
                !eps_equal(obs->elevation, OG_MISSING_R, 1.f)) { found = nx; break; }
            else if ((nx->height.data - obs->elevation > 10.f) && !eps_equal(nx->height.data, OG_MISSING_R, 1.f) &&
                     !eps_equal(obs->elevation, OG_MISSING_R, 1.f)) return 0;
        }
        if (!found) return -1;
    } else {
        int use = (obs->n_meas >= 2) ? 1 : request_p_diff;     /* :720-731 */
        for (int q = 0; q < obs->n_meas; ++q) {
            og_meas* nx = &meas[obs->first_meas + q];
            if (fabsf(nx->pressure.data - request_level * 100.f) < (float)use) { found = nx; break; }
        }
        if (!found) return -1;
    }
    switch (var) {
    case OG_VAR_UU: found->u.data = value; found->u.qc = qc; break;
    case OG_VAR_VV: found->v.data = value; found->v.qc = qc; break;
    case OG_VAR_TT: found->temperature.data = value; found->temperature.qc = qc; break;
    case OG_VAR_RH: found->rh.data = value; found->rh.qc = qc; break;
    case OG_VAR_DEWPOINT: found->dew_point.data = value; found->dew_point.qc = qc; found->rh.qc = qc; break;
    default: break;
    }
    return 0;
}

/* proc_ob_access 'get' (is_get != 0) / 'use': returns request_numobs. */
int ogo_get_obs(og_report* obs, int total_numobs, og_meas* meas, int is_get, int var, float request_level,
                int date, int time, int request_p_diff, int request_qc_max, int iew_alloc, int jns_alloc,
                float* get_value, float* get_x, float* get_y, float* get_lon, float* get_elev, int* get_index,
                int* get_qc)
{
    int request_numobs = 0, counter = 0;
    for (int loop = 0; loop < total_numobs; ++loop) {
        if (var != OG_VAR_PMSL && obs[loop].n_meas == 0) continue;     /* no_upper_air */
        if (obs[loop].discard) continue;
        if (is_get && obs[loop].bogus) continue;
        int qc = OG_MISSING_I;
        float value = 0.f;
        ogo_query_ob(&obs[loop], meas, date, time, var, request_level, request_qc_max, request_p_diff, &value, &qc);
        if (qc != OG_MISSING_I) {
            get_value[counter] = value;
            get_x[counter] = obs[loop].x; get_y[counter] = obs[loop].y;
            get_lon[counter] = obs[loop].longitude;
            get_index[counter] = loop;
            get_qc[counter] = qc;
            get_elev[counter] = obs[loop].elevation;
            if ((get_x[counter] >= 2.f) && (get_x[counter] <= (float)(iew_alloc - 1)) &&
                (get_y[counter] >= 2.f) && (get_y[counter] <= (float)(jns_alloc - 1))) {
                request_numobs = counter + 1;
                counter = counter + 1;
            } else {
                get_qc[counter] = OG_OUTSIDE_OF_DOMAIN;
                obs[loop].discard = 1;
            }
        }
    }
    return request_numobs;
}

/* qc_consistency over every measurement of every report, qc0_module.F90:313-434 */
void ogo_qc_consistency_reports(og_report* obs, int num_obs, og_meas* meas)
{
    for (int o = 0; o < num_obs; ++o)
        for (int q = 0; q < obs[o].n_meas; ++q) {
            og_meas* cur = &meas[obs[o].first_meas + q];
            int qc_u = cur->u.qc, qc_v = cur->v.qc;
            if (qc_u > qc_v) cur->v.qc = qc_u;
            if (qc_v > qc_u) cur->u.qc = qc_v;
            cur->speed.qc = cur->u.qc;
            cur->direction.qc = cur->u.qc;
            int qc_t = cur->temperature.qc, qc_rh = cur->rh.qc;
            if (ogo_contains_2n(qc_t, OG_FAILS_ERROR_MAX) && !ogo_contains_2n(qc_rh, OG_FAILS_ERROR_MAX)) {
                ogo_add_to_qc_flag(&cur->rh.qc, OG_FAILS_ERROR_MAX);
                ogo_add_to_qc_flag(&cur->dew_point.qc, OG_FAILS_ERROR_MAX);
            }
            if (ogo_contains_2n(qc_t, OG_FAILS_BUDDY_CHECK) && !ogo_contains_2n(qc_rh, OG_FAILS_BUDDY_CHECK)) {
                ogo_add_to_qc_flag(&cur->rh.qc, OG_FAILS_BUDDY_CHECK);
                ogo_add_to_qc_flag(&cur->dew_point.qc, OG_FAILS_BUDDY_CHECK);
            }
            if ((cur->dew_point.data > cur->temperature.data) && (qc_rh < OG_FAILS_BUDDY_CHECK)) {
                ogo_add_to_qc_flag(&cur->rh.qc, OG_FAILS_BUDDY_CHECK);
                ogo_add_to_qc_flag(&cur->dew_point.qc, OG_FAILS_BUDDY_CHECK);
            }
        }
}

/* proc_qc (src/proc_qc.F90:243-408) over the report list: level x variable loop of 'get', error_max,
 * buddy_check, 'put', then qc_consistency.  Fields come from `d` (t, u, v, rh, slp as og_time_desc
 * holds them); the observation rows of `d` are not used.  Returns 0, or -1 if a 'put' lost its level. */
int ogo_qc_time_reports(const og_time_desc* d, const og_qc_params* p, og_report* obs, int nrep, og_meas* meas,
                        int date, int time, int request_p_diff)
{
    const int iew = d->iew, jns = d->jns;
    const size_t npts = (size_t)iew * jns;
    size_t nn = (size_t)(nrep > 0 ? nrep : 1);
    float* gridded = (float*)malloc(sizeof(float) * npts);
    float* gt = (float*)malloc(sizeof(float) * npts);
    float* val = (float*)malloc(sizeof(float) * nn); float* x = (float*)malloc(sizeof(float) * nn);
    float* y = (float*)malloc(sizeof(float) * nn); float* lon = (float*)malloc(sizeof(float) * nn);
    float* elev = (float*)malloc(sizeof(float) * nn);
    int* qc = (int*)malloc(sizeof(int) * nn); int* idx = (int*)malloc(sizeof(int) * nn); int* lt = (int*)malloc(sizeof(int) * nn);
    int rc = 0;
    for (int k = 0; k < d->kbu; ++k) {
        const float pk = d->pressure[k];
        for (int ivar = 1; ivar <= 7; ++ivar) {
            if (k > 0 && (ivar == 5 || ivar == 6)) continue;
            if (ivar == 6) continue;
            if (ivar == 7 && !p->qc_dewpoint) continue;
            float error_difference = 0.f, buddy_difference = 0.f;
            const float* f3 = NULL;
            switch (ivar) {
            case 1: error_difference = p->max_error_uv; buddy_difference = p->max_buddy_uv; f3 = d->u; break;
            case 2: error_difference = p->max_error_uv; buddy_difference = p->max_buddy_uv; f3 = d->v; break;
            case 3: error_difference = p->max_error_t; buddy_difference = p->max_buddy_t; f3 = d->t; break;
            case 4:
                if (pk == 300.f) error_difference = p->max_error_rh * 2.f;
                else if (pk < 300.f) error_difference = p->max_error_rh * 3.f;
                else error_difference = p->max_error_rh;
                buddy_difference = p->max_buddy_rh; f3 = d->rh; break;
            case 5: error_difference = p->max_error_p; buddy_difference = p->max_buddy_p; f3 = d->slp; break;
            case 7: error_difference = p->max_error_dewpoint; buddy_difference = p->max_buddy_dewpoint; f3 = d->rh; break;
            }
            const float* yx = f3 + (ivar == 5 ? 0 : (size_t)k * npts);
            for (int i = 0; i < iew; ++i) for (int j = 0; j < jns; ++j) gridded[(size_t)j * iew + i] = yx[(size_t)i * jns + j];
            if (ivar == 5) for (size_t q = 0; q < npts; ++q) gridded[q] = gridded[q] * 100.f;
            if (ivar == 7) {
                const float* ty = d->t + (size_t)k * npts;
                for (int i = 0; i < iew; ++i) for (int j = 0; j < jns; ++j) gt[(size_t)j * iew + i] = ty[(size_t)i * jns + j];
                for (size_t q = 0; q < npts; ++q) gridded[q] = ogo_dewpoint(gt[q], gridded[q]);
            }
            int n = ogo_get_obs(obs, nrep, meas, 1, ivar, pk, date, time, request_p_diff, 1 << 20, iew, jns,
                                val, x, y, lon, elev, idx, qc);
            ogo_local_time(p->time_hhmmss / 100, lon, n, lt);
            if (p->qc_test_error_max)
                ogo_error_max(val, x, y, elev, qc, n, gridded, iew, jns, ivar, error_difference, pk, lt, NULL, NULL, NULL);
            if (p->qc_test_buddy)
                ogo_buddy_check(val, n, x, y, qc, elev, gridded, iew, jns, ivar, pk, d->dx_km, p->buddy_weight,
                                buddy_difference, NULL, NULL, NULL, NULL);
            for (int m = 0; m < n; ++m)                                     /* 'put' */
                if (ogo_store_ob(&obs[idx[m]], meas, ivar, pk, request_p_diff, val[m], qc[m]) != 0) rc = -1;
        }
    }
    ogo_qc_consistency_reports(obs, nrep, meas);
    free(gridded); free(gt); free(val); free(x); free(y); free(lon); free(elev); free(qc); free(idx); free(lt);
    return rc;
}
