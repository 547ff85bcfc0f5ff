// og_reports.cpp -- report list <-> flat per-time observation table (include/obsgrid_reports.h).
//
// Host code only.  The reference answers every (level, variable) slab's request by walking every
// report's measurement list (proc_ob_access.F90:146-379 -> query_ob, ob_access_module.F90:11-535);
// here the slab-independent part of that walk -- time window, which measurement of the list is "the
// surface" or "this pressure level" -- is done once per analysis time and the result laid out as the
// table rows the device-side selection reads (og_time.cu: sel_pred / sel_scatter_kernel apply the
// slab-dependent rest: value present, flag below request_qc_max, inside the domain, DEWPOINT).
#include "../../include/obsgrid_reports.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

namespace og { void set_error(const char* fmt, ...); }

namespace {

inline bool not_missing(float r) { return fabsf(r - OG_MISSING_R) > 1.f; }          // ob_access_module.F90:65
inline bool eps_equal(float a, float b, float eps) { return fabsf(a - b) < eps; }   // obs_sort_module.F90:304-319

// contains_2n (obs_sort_module.F90:4515-4560) for a non-negative flag word and a power of two
inline bool contains_2n(int q, int bit) { return q >= 0 && (q & bit) != 0; }

// add_to_qc_flag (qc0_module.F90:40-76): "set bit" in MOD / integer-division arithmetic
inline void add_to_qc_flag(int& q, int bit)
{
    const int qc_small = q % bit, qc_large = (q / (bit * 2)) * 2;
    q = qc_large * bit + qc_small + bit;
}

// days since 1970-01-01 of a proleptic Gregorian date (geth_newdate's calendar, date_pack_module)
long long days_from_civil(long long y, int m, int d)
{
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const int yoe = (int)(y - era * 400);
    const int doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}

// CCYYMMDDHHmmss -> seconds; false when a character is not a digit
bool seconds_of(const char* c14, long long* out)
{
    int v[14];
    for (int i = 0; i < 14; ++i) {
        if (c14[i] < '0' || c14[i] > '9') return false;
        v[i] = c14[i] - '0';
    }
    const int y = v[0] * 1000 + v[1] * 100 + v[2] * 10 + v[3], mo = v[4] * 10 + v[5], d = v[6] * 10 + v[7];
    const int h = v[8] * 10 + v[9], mi = v[10] * 10 + v[11], s = v[12] * 10 + v[13];
    *out = days_from_civil(y, mo, d) * 86400LL + h * 3600 + mi * 60 + s;
    return true;
}

bool analysis_seconds(int date, int time, long long* out)
{
    char adate[32];
    snprintf(adate, sizeof adate, "%04d%02d%02d%02d%02d%02d", date / 10000, (date / 100) % 100, date % 100,
             time / 10000, (time / 100) % 100, time % 100);
    return seconds_of(adate, out);
}

// half width of the time window in seconds, ob_access_module.F90:88-112
int window_seconds(const char* platform)
{
    auto starts = [&](const char* lit) { return strncmp(platform, lit, strlen(lit)) == 0; };
    const int by_platform = (starts("FM-97 AIREP") || starts("FM-36 TEMP SHIP") || starts("FM-35 TEMP") ||
                             starts("FM-88 SATOB")) ? 3600 : 1800;
    const char* f = platform + 35;          // platform(36:39)
    if (f[0] == ' ' && f[1] == ' ' && f[2] == ' ' && f[3] == ' ') return by_platform;
    // READ ( ... , FMT='(I4)' ): blanks are ignored, an optional sign, then digits
    int i = 0, sign = 1, val = 0, ndig = 0;
    while (i < 4 && f[i] == ' ') ++i;
    if (i < 4 && (f[i] == '+' || f[i] == '-')) { sign = f[i] == '-' ? -1 : 1; ++i; }
    for (; i < 4; ++i) {
        if (f[i] == ' ') continue;
        if (f[i] < '0' || f[i] > '9') return by_platform;      // IOSTAT /= 0
        val = val * 10 + (f[i] - '0'); ++ndig;
    }
    return ndig ? sign * val : by_platform;
}

// The measurement query_ob answers a surface request with (ob_access_module.F90:155-309), or -1
int find_surface(const og_report& R, const og_meas* m)
{
    for (int q = 0; q < R.n_meas; ++q) {
        const float h = m[R.first_meas + q].height.data;
        if (eps_equal(h, R.elevation, 1.f) && !eps_equal(h, OG_MISSING_R, 1.f) && !eps_equal(R.elevation, OG_MISSING_R, 1.f))
            return R.first_meas + q;
        if ((h - R.elevation > 10.f) && !eps_equal(h, OG_MISSING_R, 1.f) && !eps_equal(R.elevation, OG_MISSING_R, 1.f))
            return -1;          // vertically ordered: the surface is not above here
    }
    return -1;
}

// ... a pressure-level request with (:311-489), or -1; *single: the report has one level and was
// matched within request_p_diff (it then gets the extend_influence flag, :387-403)
int find_level(const og_report& R, const og_meas* m, float request_level, int request_p_diff, bool* single)
{
    *single = false;
    const float want = request_level * 100.f;
    int uselev = -1;
    float minpdiff = 9999999.f;
    for (int q = 0; q < R.n_meas; ++q) {
        const float curpdiff = fabsf(m[R.first_meas + q].pressure.data - want);
        if (curpdiff < (float)request_p_diff && curpdiff < minpdiff) { minpdiff = curpdiff; uselev = q; }
    }
    const int maxlev = R.n_meas;
    for (int q = 0; q < R.n_meas; ++q) {
        const og_meas& v = m[R.first_meas + q];
        const float curpdiff = fabsf(v.pressure.data - want);
        const bool one = (maxlev == 1) && (q == uselev);
        if ((one || ((maxlev != 1) && (curpdiff < 1.f))) &&
            (!eps_equal(v.height.data, R.elevation, 1.f) || eps_equal(v.height.data, OG_MISSING_R, 1.f))) {
            *single = one;
            return R.first_meas + q;
        }
    }
    return -1;
}

// the time window of query_ob, ob_access_module.F90:70-121
bool in_time_window(const og_report& R, long long t_an)
{
    long long t_ob = 0;
    const int ds = window_seconds(R.platform);
    return seconds_of(R.date_char, &t_ob) && (t_ob <= t_an + ds) && (t_ob >= t_an - ds);
}
// ... and what every request does to a report outside it, :511-518
void mark_no_qc_possible(const og_report& R, og_meas* meas)
{
    if (R.n_meas <= 0) return;
    og_meas& s = meas[R.first_meas];
    s.pressure.qc = s.height.qc = s.temperature.qc = s.u.qc = s.v.qc = s.rh.qc = OG_NO_QC_POSSIBLE;
}

inline bool in_domain(float x, float y, int iew, int jns)   // proc_ob_access.F90:231-234 (the 'U'/'V' branch never matches)
{
    return (x >= 2.f) && (x <= (float)(iew - 1)) && (y >= 2.f) && (y <= (float)(jns - 1));
}

int check(const og_report* reports, int nrep, const og_meas* meas, int nmeas, const og_time_desc* t)
{
    if (nrep < 0 || nmeas < 0 || (nrep > 0 && !reports) || (nmeas > 0 && !meas) || !t) {
        og::set_error("og_reports: null or negative argument"); return OG_ERR_INVALID;
    }
    if (t->nrep != nrep || t->kbu < 1 || !t->pressure) { og::set_error("og_reports: table dimensions do not match the report list"); return OG_ERR_INVALID; }
    for (int r = 0; r < nrep; ++r)
        if (reports[r].n_meas < 0 || reports[r].first_meas < 0 || reports[r].first_meas + reports[r].n_meas > nmeas) {
            og::set_error("og_reports: report %d points outside the measurement array", r); return OG_ERR_INVALID;
        }
    return OG_OK;
}

} // namespace

extern "C" int og_reports_to_table(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                                   int request_p_diff, og_time_desc* T, int* meas_index)
{
    if (int rc = check(reports, nrep, meas, nmeas, T)) return rc;
    if (nrep > 0 && (!meas_index || !T->ob_u || !T->ob_v || !T->ob_t || !T->ob_rh || !T->ob_slp || !T->qc_u ||
                     !T->qc_v || !T->qc_t || !T->qc_rh || !T->qc_slp)) {
        og::set_error("og_reports_to_table: null table row"); return OG_ERR_INVALID;
    }
    const int kbu = T->kbu;
    long long t_an = 0;
    if (!analysis_seconds(date, time, &t_an)) { og::set_error("og_reports_to_table: bad analysis date %d %d", date, time); return OG_ERR_INVALID; }

    for (int r = 0; r < nrep; ++r) {
        og_report& R = reports[r];
        // rows of a report that answers nothing: missing everywhere
        for (int k = 0; k < kbu; ++k) {
            const size_t e = (size_t)k * nrep + r;
            ((float*)T->ob_u)[e] = ((float*)T->ob_v)[e] = ((float*)T->ob_t)[e] = ((float*)T->ob_rh)[e] = OG_MISSING_R;
            T->qc_u[e] = T->qc_v[e] = T->qc_t[e] = T->qc_rh[e] = 0;
            if (T->ob_td) T->ob_td[e] = OG_MISSING_R;
            if (T->qc_td) T->qc_td[e] = 0;
            meas_index[e] = -1;
        }
        ((float*)T->ob_slp)[r] = OG_MISSING_R; T->qc_slp[r] = 0;
        if (T->xob) ((float*)T->xob)[r] = R.x;
        if (T->yob) ((float*)T->yob)[r] = R.y;
        if (T->lon) ((float*)T->lon)[r] = R.longitude;
        if (T->elev) ((float*)T->elev)[r] = R.elevation;
        if (T->bogus) ((int*)T->bogus)[r] = R.bogus ? 1 : 0;
        if (R.discard) continue;                                    // proc_ob_access.F90:171-173

        if (!in_time_window(R, t_an)) {                             // ob_access_module.F90:70-121, :511-518
            // (a bogus report is first asked by proc_oa, after qc_consistency: og_table_to_reports marks it)
            if (!R.bogus) mark_no_qc_possible(R, meas);
            continue;
        }
        ((float*)T->ob_slp)[r] = R.slp.data;                        // 'PMSL    ', :131-137
        T->qc_slp[r] = R.slp.qc;
        if (R.n_meas == 0) continue;                                // no_upper_air, proc_ob_access.F90:163-166
        // A report outside the domain is discarded by the first request that finds it
        // (proc_ob_access.F90:237-240) and never asked again: if that is the PMSL request of the
        // surface level, the pressure-level requests that would mark a single-level report with
        // extend_influence are never made.  ('get' skips bogus reports: they meet 'use' first.)
        const int first_qmax = R.bogus ? (OG_FAILS_ERROR_MAX < OG_FAILS_BUDDY_CHECK ? OG_FAILS_ERROR_MAX : OG_FAILS_BUDDY_CHECK) : (1 << 20);
        const bool gone_at_pmsl = !in_domain(R.x, R.y, T->iew, T->jns) && not_missing(R.slp.data) && (R.slp.qc < first_qmax);
        for (int k = 0; k < kbu; ++k) {
            const float p = T->pressure[k];
            bool single = false;
            const int mi = ((int)lroundf(p) == 1001) ? find_surface(R, meas) : find_level(R, meas, p, request_p_diff, &single);
            if (mi < 0) continue;
            og_meas& v = meas[mi];
            if (single && !gone_at_pmsl && !contains_2n(v.pressure.qc, OG_EXTEND_INFLUENCE)) v.pressure.qc += OG_EXTEND_INFLUENCE;
            const size_t e = (size_t)k * nrep + r;
            ((float*)T->ob_u)[e] = v.u.data;            T->qc_u[e] = v.u.qc;
            ((float*)T->ob_v)[e] = v.v.data;            T->qc_v[e] = v.v.qc;
            ((float*)T->ob_t)[e] = v.temperature.data;  T->qc_t[e] = v.temperature.qc;
            ((float*)T->ob_rh)[e] = v.rh.data;          T->qc_rh[e] = v.rh.qc;
            if (T->ob_td) T->ob_td[e] = v.dew_point.data;
            if (T->qc_td) T->qc_td[e] = v.dew_point.qc;
            meas_index[e] = mi;
        }
    }
    return OG_OK;
}

extern "C" int og_table_shared_levels(const int* meas_index, int kbu, int nrep)
{
    if (!meas_index || kbu < 1 || nrep < 0) return 0;
    int shared = 0;
    for (int r = 0; r < nrep; ++r)
        for (int k = 1; k < kbu; ++k) {
            const int mi = meas_index[(size_t)k * nrep + r];
            if (mi < 0) continue;
            for (int kk = 0; kk < k; ++kk)
                if (meas_index[(size_t)kk * nrep + r] == mi) { ++shared; break; }
        }
    return shared;
}

extern "C" int og_table_to_reports(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                                   const og_time_desc* T, const int* meas_index)
{
    if (int rc = check(reports, nrep, meas, nmeas, T)) return rc;
    long long t_an = 0;
    if (!analysis_seconds(date, time, &t_an)) { og::set_error("og_table_to_reports: bad analysis date %d %d", date, time); return OG_ERR_INVALID; }
    if (nrep > 0 && (!meas_index || !T->qc_u || !T->qc_v || !T->qc_t || !T->qc_rh || !T->qc_slp || !T->ob_slp)) {
        og::set_error("og_table_to_reports: null table row"); return OG_ERR_INVALID;
    }
    const int kbu = T->kbu;
    for (int r = 0; r < nrep; ++r) {
        og_report& R = reports[r];
        if (R.discard) continue;
        const float x = T->xob ? T->xob[r] : R.x, y = T->yob ? T->yob[r] : R.y;
        // something a 'get' (flag below 2**20; bogus reports are skipped) or a 'use' (below the first
        // failed-check bit) would have found?
        const int qmax = R.bogus ? (OG_FAILS_ERROR_MAX < OG_FAILS_BUDDY_CHECK ? OG_FAILS_ERROR_MAX : OG_FAILS_BUDDY_CHECK) : (1 << 20);
        bool found = not_missing(T->ob_slp[r]) && T->qc_slp[r] < qmax;
        for (int k = 0; k < kbu && !found; ++k) {
            const size_t e = (size_t)k * nrep + r;
            if (meas_index[e] < 0) continue;
            found = (not_missing(T->ob_u[e]) && T->qc_u[e] < qmax) || (not_missing(T->ob_v[e]) && T->qc_v[e] < qmax) ||
                    (not_missing(T->ob_t[e]) && T->qc_t[e] < qmax) || (not_missing(T->ob_rh[e]) && T->qc_rh[e] < qmax);
        }
        if (!found) continue;
        if (!in_domain(x, y, T->iew, T->jns)) { R.discard = 1; continue; }     // proc_ob_access.F90:237-240
        if (not_missing(T->ob_slp[r])) R.slp.qc = T->qc_slp[r];                 // store_ob :572-574
        for (int k = 0; k < kbu; ++k) {
            const size_t e = (size_t)k * nrep + r;
            const int mi = meas_index[e];
            if (mi < 0) continue;
            og_meas& v = meas[mi];
            // With a pressure tolerance wider than half the spacing of the analysis levels one
            // single-level measurement answers two of them.  The reference checks it at the lower
            // level, stores the flags, and checks it again at the next level starting from those
            // flags; the checks only set bits and do not read them, so the measurement ends with
            // the union of what its levels set -- here: OR into what an earlier level left.
            bool again = false;
            for (int kk = 0; kk < k && !again; ++kk) again = meas_index[(size_t)kk * nrep + r] == mi;
            // (store_ob looks for the level on its own, ob_access_module.F90:708-740, and does not skip
            // the surface measurement as query_ob does, :399-402: a sounding whose surface pressure is
            // within 1 Pa of an analysis level it also reports has that level's checked value and flag
            // stored into its SURFACE measurement by the reference.  That defect is not reproduced:
            // flags go back to the measurement the value was read from.  include/obsgrid_reports.h)
            auto put = [&](og_field& f, int q) { f.qc = again ? (f.qc | q) : q; };
            if (not_missing(T->ob_u[e])) put(v.u, T->qc_u[e]);
            if (not_missing(T->ob_v[e])) put(v.v, T->qc_v[e]);
            if (not_missing(T->ob_t[e])) put(v.temperature, T->qc_t[e]);
            if (not_missing(T->ob_rh[e])) put(v.rh, T->qc_rh[e]);
            if (T->ob_td && T->qc_td) {                                         // store_ob 'DEWPOINT', :645-657 / :771-774
                v.dew_point.data = T->ob_td[e];
                put(v.dew_point, T->qc_td[e]);
            }
        }
    }
    // qc_consistency (qc0_module.F90:313-434) over every measurement of every report: the table rows
    // have been through it on the device (it is idempotent); this covers the levels of a report that
    // are no analysis level, and the speed / direction copies of the wind flag
    for (int r = 0; r < nrep; ++r)
        for (int q = 0; q < reports[r].n_meas; ++q) {
            og_meas& c = meas[reports[r].first_meas + q];
            const int qc_u = c.u.qc, qc_v = c.v.qc;
            if (qc_u > qc_v) c.v.qc = qc_u;
            if (qc_v > qc_u) c.u.qc = qc_v;
            c.speed.qc = c.u.qc;
            c.direction.qc = c.u.qc;
            const int qc_t = c.temperature.qc, qc_rh = c.rh.qc;
            if (contains_2n(qc_t, OG_FAILS_ERROR_MAX) && !contains_2n(qc_rh, OG_FAILS_ERROR_MAX)) {
                add_to_qc_flag(c.rh.qc, OG_FAILS_ERROR_MAX); add_to_qc_flag(c.dew_point.qc, OG_FAILS_ERROR_MAX);
            }
            if (contains_2n(qc_t, OG_FAILS_BUDDY_CHECK) && !contains_2n(qc_rh, OG_FAILS_BUDDY_CHECK)) {
                add_to_qc_flag(c.rh.qc, OG_FAILS_BUDDY_CHECK); add_to_qc_flag(c.dew_point.qc, OG_FAILS_BUDDY_CHECK);
            }
            if ((c.dew_point.data > c.temperature.data) && (qc_rh < OG_FAILS_BUDDY_CHECK)) {
                add_to_qc_flag(c.rh.qc, OG_FAILS_BUDDY_CHECK); add_to_qc_flag(c.dew_point.qc, OG_FAILS_BUDDY_CHECK);
            }
        }
    // proc_oa's 'use' requests come after qc_consistency and meet the reports outside the time
    // window once more (query_ob :511-518 assigns, it does not add)
    for (int r = 0; r < nrep; ++r)
        if (!reports[r].discard && !in_time_window(reports[r], t_an)) mark_no_qc_possible(reports[r], meas);
    return OG_OK;
}
