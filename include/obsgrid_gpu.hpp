// obsgrid_gpu.hpp -- C++ host-side mirror of the reference interface over the C ABI.
//
// OBSGRID is compiled code (Fortran 90); the image this library is developed in has no Fortran
// compiler, so the compiled host side that is exercised here is this header (C++17, header only)
// with host/obsgrid_host.cpp as its driver -- the same shape as fortran/obsgrid_gpu_iface.F90,
// which forwards the reference's own procedures to the same entry points.  Procedure names,
// argument order and meaning are the reference's:
//
//   cressman      src/oa_module.F90:25-32        error_max    src/qc1_module.F90:9-15
//   buddy_check   src/qc2_module.F90:9-14        proc_qc      src/proc_qc.F90:30-33 (level x variable loop)
//   proc_oa       src/proc_oa.F90:30-40          mixing_ratio src/final_analysis_module.F90:1026
//
// Errors: the C functions return a negative og_status and never abort; like the reference's
// error_handler(..., fatal=.TRUE.) (src/error_handler.F90:53-92) the wrappers do not continue --
// they throw og::Error carrying og_last_error().  There is no CPU path: constructing a Context
// without an sm_100 device throws.
#pragma once
#include <stdexcept>
#include <string>

#include "obsgrid_gpu.h"
#include "obsgrid_reports.h"

namespace og {

struct Error : std::runtime_error {
    int status;
    Error(const std::string& what, int rc)
        : std::runtime_error(what + " failed with og_status " + std::to_string(rc) + ": " +
                             (og_last_error() ? og_last_error() : "")),
          status(rc) {}
};

class Context {
public:
    explicit Context(int device = 0)
    {
        const int rc = og_init(&h_, device);
        if (rc != OG_OK) throw Error("og_init", rc);
    }
    ~Context() { if (h_) og_finalize(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    og_ctx* handle() const { return h_; }

    void set_arithmetic(int mode) { check(og_set_arithmetic(h_, mode), "og_set_arithmetic"); }

    // One slab over several GPUs: installs the all-reduce callback (obsgrid_gpu.h, og_set_exchange);
    // every rank must then make the same calls on the same batches.  world <= 1: back to one GPU.
    void set_exchange(int rank, int world, og_exchange_fn fn, void* user)
    {
        check(og_set_exchange(h_, rank, world, fn, user), "og_set_exchange");
    }

    // The report list of one analysis time -> the flat table proc_qc / proc_oa run on, and the flags
    // back (obsgrid_reports.h: query_ob / store_ob / proc_ob_access / qc_consistency semantics).
    static void reports_to_table(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                                 int request_p_diff, og_time_desc& table, int* meas_index)
    {
        check(og_reports_to_table(reports, nrep, meas, nmeas, date, time, request_p_diff, &table, meas_index),
              "og_reports_to_table");
    }
    static void table_to_reports(og_report* reports, int nrep, og_meas* meas, int nmeas, int date, int time,
                                 const og_time_desc& table, const int* meas_index)
    {
        check(og_table_to_reports(reports, nrep, meas, nmeas, date, time, &table, meas_index), "og_table_to_reports");
    }

    // SUBROUTINE cressman ( obs_value , obs , xob , yob , numobs , gridded , iew , jns , crsdot , name ,
    //   radius_influence , dxd , u_banana , v_banana , pressure , passes , smooth_type , lat_center ,
    //   use_first_guess , fg_value , scale_cressman_rh_decreases )      -- obs_value, lat_center unused
    void cressman(const float* /*obs_value*/, const float* obs, const float* xob, const float* yob,
                  int numobs, float* gridded, int iew, int jns, int crsdot, int name /*OG_VAR_*/,
                  int radius_influence, float dxd, const float* u_banana, const float* v_banana,
                  float pressure, int passes, int smooth_type, float /*lat_center*/,
                  bool use_first_guess, const float* fg_value, bool scale_cressman_rh_decreases) const
    {
        check(og_cressman(h_, obs, xob, yob, numobs, gridded, iew, jns, crsdot, name, radius_influence,
                          dxd, u_banana, v_banana, pressure, passes, smooth_type, use_first_guess ? 1 : 0,
                          fg_value, scale_cressman_rh_decreases ? 1 : 0), "cressman");
    }

    // SUBROUTINE error_max ( station_id , obs , ship_report , xob , yob , station_elevation , index ,
    //   qc_flag , numobs , gridded , iew , jns , name , max_difference , pressure , local_time ,
    //   print_error_max )                -- ids, ship_report, index, print: host-side only
    void error_max(const float* obs, const float* xob, const float* yob, const float* station_elevation,
                   int* qc_flag, int numobs, const float* gridded, int iew, int jns, int name,
                   float max_difference, float pressure, const int* local_time) const
    {
        check(og_error_max(h_, obs, xob, yob, station_elevation, qc_flag, numobs, gridded, iew, jns, name,
                           max_difference, pressure, local_time, nullptr, nullptr, nullptr), "error_max");
    }

    // SUBROUTINE buddy_check ( station_id , obs , numobs , xob , yob , qc_flag , station_elevation ,
    //   gridded , iew , jns , name , pressure , local_time , dx , buddy_weight , buddy_difference ,
    //   print_buddy )
    void buddy_check(const float* obs, int numobs, const float* xob, const float* yob, int* qc_flag,
                     const float* station_elevation, const float* gridded, int iew, int jns, int name,
                     float pressure, float dx, float buddy_weight, float buddy_difference) const
    {
        check(og_buddy_check(h_, obs, numobs, xob, yob, qc_flag, station_elevation, gridded, iew, jns,
                             name, pressure, dx, buddy_weight, buddy_difference, nullptr, nullptr,
                             nullptr, nullptr), "buddy_check");
    }

    // proc_qc / proc_oa of one analysis time over the flat observation table
    void proc_qc(const og_time_desc& d, const og_qc_params& p) const { check(og_qc_time(h_, &d, &p), "proc_qc"); }
    void proc_qc_fdda(const og_time_desc& d, const og_qc_params& p, float* tobbox, float* odis) const
    {
        check(og_qc_time_fdda(h_, &d, &p, tobbox, odis), "proc_qc (fdda)");
    }
    // returns false where the reference STOPs on a surface radius below 4.5 cells (OG_ERR_RADIUS)
    bool proc_oa(const og_time_desc& d, const og_oa_params& p) const
    {
        const int rc = og_oa_time(h_, &d, &p, 0, d.kbu);
        if (rc == OG_ERR_RADIUS) return false;
        check(rc, "proc_oa");
        return true;
    }
    void proc_qc_oa(const og_time_desc& d, const og_qc_params& q, const og_oa_params& o) const
    {
        const int rc = og_qc_oa_time(h_, &d, &q, &o, 0, d.kbu, 0);
        if (rc != OG_ERR_RADIUS) check(rc, "proc_qc + proc_oa");
    }

    void mixing_ratio(float* rh, const float* t, const float* h, const float* terrain, const float* slp_x,
                      const float* pressure, int iew_alloc, int jns_alloc, int kbu_alloc, float* qv,
                      float* psfc) const
    {
        check(og_mixing_ratio(h_, rh, t, h, terrain, slp_x, pressure, iew_alloc, jns_alloc, kbu_alloc, qv, psfc),
              "mixing_ratio");
    }

private:
    static void check(int rc, const char* what) { if (rc != OG_OK) throw Error(what, rc); }
    og_ctx* h_ = nullptr;
};

} // namespace og
