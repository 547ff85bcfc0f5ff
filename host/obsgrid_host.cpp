// obsgrid_host.cpp -- compiled host-side driver over include/obsgrid_gpu.hpp.
//
// Plays the part of the reference's time loop body (src/driver.F90:440-500: CALL proc_qc,
// CALL proc_oa) for one analysis time read from a flat binary case file -- the form in which
// obsgrid_b200/synth.py writes the synthetic met_em first guess + little_r observation table, so
// that the compiled path does not depend on the Fortran ingest:
//
//   obsgrid_host <case.bin> <result.bin> [--contracted] [--one-call]
//
// Exit codes: 0 ok; 2 usage / file error; 3 the GPU library refused (no sm_100 device: there is
// no CPU fallback) or failed, message on stderr; 5 OG_ERR_RADIUS (the reference STOPs,
// src/proc_oa.F90:694-703) -- results are still written.
//
// File layout (little endian): "OGC1", int32 iew, jns, kbu, nrep, float dx_km, og_qc_params,
// og_oa_params (the structs of include/obsgrid_gpu.h, raw), then float pressure[kbu],
// t u v rh [kbu*iew*jns], slp [iew*jns], xob yob lon elev [nrep], ob_u ob_v ob_t ob_rh [kbu*nrep],
// ob_slp [nrep], int32 qc_u qc_v qc_t qc_rh [kbu*nrep], qc_slp [nrep].  The result file holds
// t u v rh slp qc_u qc_v qc_t qc_rh qc_slp in the same order.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "obsgrid_gpu.hpp"

namespace {

template <class T> bool rd(FILE* f, std::vector<T>& v, size_t n)
{
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
template <class T> bool wr(FILE* f, const std::vector<T>& v)
{
    return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <case.bin> <result.bin> [--contracted] [--one-call]\n", argv[0]); return 2; }
    bool contracted = false, one_call = false;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--contracted")) contracted = true;
        else if (!strcmp(argv[i], "--one-call")) one_call = true;
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[4];
    int dims[4];
    float dx_km;
    og_qc_params qp;
    og_oa_params op;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "OGC1", 4) || fread(dims, 4, 4, f) != 4 ||
        fread(&dx_km, 4, 1, f) != 1 || fread(&qp, sizeof qp, 1, f) != 1 || fread(&op, sizeof op, 1, f) != 1) {
        fprintf(stderr, "%s: not a case file\n", argv[1]); return 2;
    }
    const int iew = dims[0], jns = dims[1], kbu = dims[2], nrep = dims[3];
    const size_t per = (size_t)iew * jns, n3 = per * kbu, nt = (size_t)kbu * nrep;
    std::vector<float> pressure, t, u, v, rh, slp, xob, yob, lon, elev, ob_u, ob_v, ob_t, ob_rh, ob_slp;
    std::vector<int> qc_u, qc_v, qc_t, qc_rh, qc_slp;
    const bool ok = rd(f, pressure, (size_t)kbu) && rd(f, t, n3) && rd(f, u, n3) && rd(f, v, n3) && rd(f, rh, n3) &&
                    rd(f, slp, per) && rd(f, xob, (size_t)nrep) && rd(f, yob, (size_t)nrep) && rd(f, lon, (size_t)nrep) &&
                    rd(f, elev, (size_t)nrep) && rd(f, ob_u, nt) && rd(f, ob_v, nt) && rd(f, ob_t, nt) && rd(f, ob_rh, nt) &&
                    rd(f, ob_slp, (size_t)nrep) && rd(f, qc_u, nt) && rd(f, qc_v, nt) && rd(f, qc_t, nt) &&
                    rd(f, qc_rh, nt) && rd(f, qc_slp, (size_t)nrep);
    fclose(f);
    if (!ok) { fprintf(stderr, "%s: truncated case file\n", argv[1]); return 2; }

    og_time_desc d{};
    d.iew = iew; d.jns = jns; d.kbu = kbu; d.dx_km = dx_km; d.pressure = pressure.data();
    d.t = t.data(); d.u = u.data(); d.v = v.data(); d.rh = rh.data(); d.slp = slp.data();
    d.nrep = nrep; d.xob = xob.data(); d.yob = yob.data(); d.lon = lon.data(); d.elev = elev.data();
    d.ob_u = ob_u.data(); d.ob_v = ob_v.data(); d.ob_t = ob_t.data(); d.ob_rh = ob_rh.data(); d.ob_slp = ob_slp.data();
    d.qc_u = qc_u.data(); d.qc_v = qc_v.data(); d.qc_t = qc_t.data(); d.qc_rh = qc_rh.data(); d.qc_slp = qc_slp.data();

    int rc = 0;
    try {
        og::Context ctx(0);
        if (contracted) ctx.set_arithmetic(OG_ARITH_CONTRACTED);
        if (one_call) {
            ctx.proc_qc_oa(d, qp, op);
        } else {
            ctx.proc_qc(d, qp);                       // driver.F90:456-459
            if (!ctx.proc_oa(d, op)) rc = 5;          // driver.F90:489-492
        }
    } catch (const og::Error& e) {
        fprintf(stderr, "obsgrid_host: %s\n", e.what());
        return 3;
    }
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    const bool wok = wr(o, t) && wr(o, u) && wr(o, v) && wr(o, rh) && wr(o, slp) && wr(o, qc_u) && wr(o, qc_v) &&
                     wr(o, qc_t) && wr(o, qc_rh) && wr(o, qc_slp);
    fclose(o);
    if (!wok) { fprintf(stderr, "%s: write failed\n", argv[2]); return 2; }
    return rc;
}
