/*
 * obsgrid_reports.h -- from the reference's report list to the flat per-time observation table of
 * og_time_desc, and back (part of the C ABI of libobsgrid_gpu.so; host code only).
 *
 * The batched entry points (og_qc_time, og_oa_time, og_time_*) read one table per analysis time:
 * per level and report a value (or missing) and a QC flag.  In the reference that information sits
 * in TYPE(report) records with a linked list of TYPE(measurement) (src/obs_sort_module.F90:75-245)
 * and every (level, variable) slab re-derives its observations with a call of proc_ob_access
 * (src/proc_ob_access.F90:146-379 'get' / 'use', :381-403 'put'), which asks query_ob
 * (src/ob_access_module.F90:11-535) report by report and writes back through store_ob (:538-817).
 *
 *   og_reports_to_table   performs, once per analysis time, every decision those per-slab calls
 *                         make that does not depend on the slab: the time window (:70-121), the
 *                         level of the measurement list that answers a request (surface: :147-309,
 *                         pressure levels: :311-489), and lays the measurement's u, v, t, rh, dew
 *                         point and flags out as table rows.  What does depend on the slab -- value
 *                         not missing, flag below request_qc_max, the DEWPOINT / PMSLPSFC
 *                         derivations, inside the domain -- is evaluated on the device by the
 *                         selection kernels, slab by slab, from those rows.
 *   og_table_to_reports   is 'put' for every slab at once: flags (and the dew point QC derived
 *                         from T and RH, store_ob :645-657) go back to the measurements they came
 *                         from; reports the selection found outside the domain are marked discard
 *                         (proc_ob_access.F90:237-240).
 *
 * The Fortran side flattens its list into these plain arrays (fortran/obsgrid_gpu_iface.F90:
 * og_flatten_reports / og_unflatten_flags); the layouts below are BIND(C)-compatible.
 */
#ifndef OBSGRID_REPORTS_H
#define OBSGRID_REPORTS_H

#include "obsgrid_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* TYPE(field): a value with its QC flag (obs_sort_module.F90:147-160) */
typedef struct og_field { float data; int qc; } og_field;

/* TYPE(meas_data): one level of a report (obs_sort_module.F90:206-226) */
typedef struct og_meas {
    og_field pressure, height, temperature, dew_point, speed, direction, u, v, rh, thickness;
} og_meas;

/* What query_ob / proc_ob_access read of a TYPE(report) (obs_sort_module.F90:240-250) */
typedef struct og_report {
    float latitude, longitude;      /* location                                               */
    float x, y;                     /* latlon_to_ij(projd, latitude, longitude): computed by the
                                       caller, the map projection stays on the Fortran side     */
    float elevation;                /* info%elevation                                          */
    char  platform[40];             /* info%platform, blank padded                             */
    char  date_char[14];            /* valid_time%date_char, CCYYMMDDHHmmss                    */
    int   discard, bogus;           /* info%discard, info%bogus (0 / 1)                        */
    og_field slp;                   /* ground%slp                                              */
    og_field psfc;                  /* ground%psfc (only its qc is written, store_ob :631)     */
    int   first_meas, n_meas;       /* the `surface` linked list, flattened: meas[first_meas ..
                                       first_meas + n_meas); n_meas == 0: not ASSOCIATED       */
} og_report;

#define OG_NO_QC_POSSIBLE   (1 << 15)    /* missing.inc:35 */
#define OG_EXTEND_INFLUENCE (1 << 2)     /* missing.inc:25 */

/* Table rows this header adds to og_time_desc (all [kbu * nrep] like ob_t; nullable there):
 * see og_time_desc.ob_td / qc_td / bogus in obsgrid_gpu.h. */

/* One analysis time: fills the observation rows of `table` (caller-allocated: xob, yob, lon, elev
 * [nrep]; ob_u, ob_v, ob_t, ob_rh, ob_td, qc_u, qc_v, qc_t, qc_rh, qc_td [kbu * nrep]; ob_slp,
 * qc_slp, bogus [nrep]) from the report list.  table->iew, jns, kbu, pressure and nrep must be
 * set (nrep == the number of reports); the field pointers are not touched.
 *
 *   date, time        the analysis time as proc_qc / proc_oa pass it (CCYYMMDD, HHmmss)
 *   request_p_diff    the pressure tolerance in Pa of proc_qc.F90:187-195 / proc_oa.F90:160-168
 *                     (1 unless use_p_tolerance_one_lev)
 *   meas_index        [kbu * nrep] out: the measurement (index into meas) each table entry was
 *                     taken from, -1 where the report has no such level -- og_table_to_reports
 *                     needs it; k == 0 is the surface pseudo-level (1001 hPa)
 *
 * Side effects on the list, as query_ob has them: a report outside the time window gets
 * no_qc_possible on the flags of its first measurement (:513-518); a single-level report matched
 * within request_p_diff gets extend_influence added to its pressure flag (:387-403).
 * Reports with info%discard set contribute nothing (proc_ob_access.F90:171-173). */
int og_reports_to_table(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                        int request_p_diff, og_time_desc* table, int* meas_index);

/* How many table entries share their measurement with an entry of a lower level of the same report:
 * with use_p_tolerance_one_lev and a tolerance wider than half the spacing of two analysis levels, one
 * single-level measurement answers both.  0 -- always so with the namelist defaults (tolerance off, or
 * 700 Pa against levels at least 2500 Pa apart) -- means the table path reproduces the per-slab calls
 * of the reference exactly.  Otherwise the reference checks such a measurement level after level, each
 * time starting from the flags the previous level stored, and qc_consistency runs once at the end on
 * what the levels left.  The exact path for that regime defers the consistency pass to the list:
 *     og_qc_time_levels(ctx, table, qc, 0, kbu, run_consistency = 0);
 *     og_table_to_reports(...);      -- union of the bits the levels set, then qc_consistency
 *     og_reports_to_table(...);      -- the rows proc_oa's 'use' requests would read
 *     og_oa_time(ctx, table, oa, 0, kbu);
 * (76 of 76 such cases of profiles/exp/fuzz_reports.py identical to the as-written calls.)  With
 * the per-row consistency of og_qc_time left on, the flags that come back are still the union, but
 * qc_consistency's assignments (v.qc = qc_u) have then seen each level on its own: 1 of the 76 cases
 * differed, and the analysis of one of the two levels does not see the other level's verdict. */
int og_table_shared_levels(const int* meas_index, int kbu, int nrep);

/* 'put' of every slab: the flag rows of `table` (after og_qc_time / og_time_download) go back into
 * the measurements (u, v, temperature, rh flags; dew_point data and flag where the DEWPOINT slab
 * ran), qc_slp into ground%slp.  A report that has table entries but lies outside the domain of the
 * analysis (proc_ob_access.F90:231-240) is marked discard, its flags are left alone.  Flags go back
 * to the measurement the value was read from.  (One defect of the reference is deliberately not
 * reproduced: store_ob searches the level on its own, ob_access_module.F90:708-740, and unlike
 * query_ob, :399-402, does not skip the surface measurement -- for a sounding whose surface pressure
 * lies within 1 Pa of an analysis level that it also reports, the reference stores that level's
 * checked value and flag into the SURFACE measurement, overwriting the surface observation.)  Then
 * qc_consistency (qc0_module.F90:313-434) runs over every measurement of every report -- the table
 * rows have been through it on the device and it is idempotent; this covers the levels that are
 * no analysis level and the speed / direction copies of the wind flag -- and reports outside the
 * time window get their no_qc_possible marks once more, as proc_oa's requests leave them.
 * The list is then in the state the reference's proc_qc + qc_consistency + proc_oa leave it in. */
int og_table_to_reports(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                        const og_time_desc* table, const int* meas_index);

#ifdef __cplusplus
}
#endif
#endif /* OBSGRID_REPORTS_H */
