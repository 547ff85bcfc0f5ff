"""A second, independent restatement of the two hot loops -- numpy, written from the Fortran
source (not from oracle/og_oracle.c) -- used only to cross-check the C oracle
(tests/test_oracle_vs_numpy.py).  Every scalar is np.float32 so that each operation is one
IEEE-754 binary32 operation in source order; windows are vectorised per observation, which
does not change any result (points are independent; obs are applied in index order).

    cressman     src/oa_module.F90:25-218 (+ dist_uu_vv_rh :229-574, dist_not_uu_vv_rh :577-627)
    error_max    src/qc1_module.F90:102-249;  ob_density / obs_distance  src/qc0_module.F90:81-232
    buddy_check  src/qc2_module.F90:111-337 (+ modify_qc_flag / add_to_qc_flag, qc0_module.F90:9-76)
"""
import math

import numpy as np

F = np.float32
NO_BUDDIES, FAILS_BUDDY = 1 << 14, 1 << 17


def nint(x):          # Fortran NINT: half away from zero
    x = float(x)
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def fint(x):          # Fortran INT / real -> integer assignment: truncate
    return int(float(x))


def add_flag(q, t):   # qc0_module.F90:27-31 / :55-72
    small = q % t if q >= 0 else -((-q) % t)
    large = int(q / (t * 2)) * 2
    return large * t + small + t


def cressman(obs, xob, yob, gridded, name, R, dxd, u_b=None, v_b=None, pressure=1001.0,
             use_first_guess=True, fg_value=None, scale_rh=False, crsdot=1, stats=None):
    """gridded: (jns, iew) C-order array == Fortran gridded(iew,jns); returns the analysed copy."""
    g = np.array(gridded, np.float32)
    jns, iew = g.shape
    num = np.zeros_like(g); den = np.zeros_like(g)
    rc2 = F(crsdot) / F(2.0)
    r2 = F(R * R)
    pi = F(3.14159265358); twopi = F(2.0) * F(3.14159265358)
    dxd = F(dxd); pressure = F(pressure)
    shapes = [0, 0, 0]
    for n in range(len(obs)):
        x, y, ob = F(xob[n]), F(yob[n]), F(obs[n])
        if x <= F(1) + rc2 or y <= F(1) + rc2 or x >= F(iew) - rc2 or y >= F(jns) - rc2:
            continue
        i0 = max(nint(x - rc2) - 4 * R - 1, 1); j0 = max(nint(y - rc2) - 4 * R - 1, 1)
        i1 = min(nint(x - rc2) + 4 * R + 1, iew - crsdot); j1 = min(nint(y - rc2) + 4 * R + 1, jns - crsdot)
        if i1 < i0 or j1 < j0:
            continue
        ri = np.arange(i0, i1 + 1, dtype=np.float32)[None, :]
        rj = np.arange(j0, j1 + 1, dtype=np.float32)[:, None]
        sl = (slice(j0 - 1, j1), slice(i0 - 1, i1))
        kind = "circle"
        if name in ("UU", "VV", "RH"):
            vcrit = F(15.0) if pressure < F(500.0) else F(25.0) - F(0.02) * pressure
            u_ob = F(u_b[nint(y) - 1, nint(x) - 1]); v_ob = F(v_b[nint(y) - 1, nint(x) - 1])
            speed = np.sqrt(u_ob * u_ob + v_ob * v_ob)
            if speed >= vcrit:
                dydx = v_ob / u_ob if abs(u_ob) >= F(1.0e-5) else F(1000.0)
                x1 = x + F(3.0) * (u_ob / speed); y1 = y + F(3.0) * (v_ob / speed)
                x2 = x - F(3.0) * (u_ob / speed); y2 = y - F(3.0) * (v_ob / speed)
                a1 = max(min(nint(x1), iew), 1); b1 = max(min(nint(y1), jns), 1)
                a2 = max(min(nint(x2), iew), 1); b2 = max(min(nint(y2), jns), 1)
                d12 = np.sqrt((F(a1) - F(a2)) * (F(a1) - F(a2)) + (F(b1) - F(b2)) * (F(b1) - F(b2)))
                with np.errstate(divide="ignore", invalid="ignore"):
                    dudx = (F(u_b[b1 - 1, a1 - 1]) - F(u_b[b2 - 1, a2 - 1])) / (d12 * dxd)
                    dvdx = (F(v_b[b1 - 1, a1 - 1]) - F(v_b[b2 - 1, a2 - 1])) / (d12 * dxd)
                    numc = abs(u_ob * dvdx - v_ob * dudx) / (u_ob * u_ob)
                s = np.sqrt(F(1.0) + dydx * dydx)
                denc = (s * s) * s
                if abs(u_ob) < F(1.0e-5) or abs(v_ob) < F(1.0e-5) or abs(numc) < F(1.0e-5):
                    roc = F(1000.0)
                else:
                    roc = denc / numc
                elong = F(1.0) + F(0.7778) / vcrit * speed
                if abs(roc) < (F(3.0) * F(R)) * dxd:
                    kind = "banana"
                    ia = max(nint(x) - 3, 1); ja = fint(y); ib = fint(x); jb = max(nint(y) - 3, 1)
                    ic = min(nint(x) + 3, iew); jc = fint(y); idd = fint(x); jd = min(nint(y) + 3, jns)
                    vor = (F(v_b[jc - 1, ic - 1]) - F(v_b[ja - 1, ia - 1])) - (F(u_b[jd - 1, idd - 1]) - F(u_b[jb - 1, ib - 1]))
                    roc = roc * F(math.copysign(1.0, float(vor)))
                    i_c = fint(F(nint(x)) - roc * v_ob / speed)
                    j_c = fint(F(nint(y)) + roc * u_ob / speed)
                    theta_ob = np.arctan2(F(j_c) - y, F(i_c) - x)
                else:
                    kind = "ellipse"
        shapes[("circle", "ellipse", "banana").index(kind)] += 1
        if kind == "circle":
            d2 = ((x - rc2 - ri) * (x - rc2 - ri)) + ((y - rc2 - rj) * (y - rc2 - rj))
        elif kind == "banana":
            theta = np.arctan2(F(j_c) - rj + F(0) * ri, F(i_c) - ri + F(0) * rj).astype(np.float32)
            rij = np.sqrt((F(i_c) - ri) * (F(i_c) - ri) + (F(j_c) - rj) * (F(j_c) - rj))
            td = (theta_ob - theta).astype(np.float32)
            inv2pi = F(1.0) / twopi; invpi = F(1.0) / pi
            big = td > twopi
            td = np.where(big, td - np.trunc(td * inv2pi).astype(np.float32) * twopi, td).astype(np.float32)
            sml = td < -twopi
            td = np.where(sml, td + np.trunc(td * inv2pi).astype(np.float32) * twopi, td).astype(np.float32)
            td = np.where(td > pi, td - twopi, td).astype(np.float32)
            td = np.where(td < -pi, td + twopi, td).astype(np.float32)
            xx = roc * td * invpi
            yy = abs(roc) - rij
            d2 = xx * xx / elong + yy * yy
        else:
            xx = (u_ob * (ri - x) + v_ob * (rj - y)) / speed
            yy = (v_ob * (ri - x) - u_ob * (rj - y)) / speed
            d2 = xx * xx / elong + yy * yy
        d2 = d2.astype(np.float32)
        hit = d2 < r2
        if stats is not None:
            st = stats.setdefault(kind, [0, 0, 0])
            st[0] += 1; st[1] += hit.size; st[2] += int(np.count_nonzero(hit))
        w = ((r2 - d2) / (r2 + d2)).astype(np.float32)
        sf = np.ones_like(w)
        if scale_rh and name == "RH" and use_first_guess and ob < F(0.0):
            with np.errstate(divide="ignore", invalid="ignore"):
                sf = np.minimum(F(1.0), g[sl] / F(fg_value[n])).astype(np.float32)
        num[sl] = np.where(hit, num[sl] + w * w * ob * sf, num[sl])
        den[sl] = np.where(hit, den[sl] + w, den[sl])
    pert = np.zeros_like(g)
    pos = den > 0
    pert[pos] = num[pos] / den[pos]
    out = (g + pert) if use_first_guess else pert
    return out.astype(np.float32), shapes


def bilinear(g, x, y):
    iob, job = fint(x), fint(y)
    dx, dy = F(x) - F(iob), F(y) - F(job)
    G = lambda i, j: F(g[j - 1, i - 1])   # noqa: E731
    return (F(1) - dx) * ((F(1) - dy) * G(iob, job) + dy * G(iob, job + 1)) + \
        dx * ((F(1) - dy) * G(iob + 1, job) + dy * G(iob + 1, job + 1))


def buddy_check(obs, xob, yob, qc_flag, elev, gridded, name, pressure, dx, buddy_weight, buddy_difference):
    n = len(obs)
    p = F(pressure); bd = F(buddy_difference)
    rng = F(650.0) if p <= F(201.0) else (F(300.0) if p <= F(1000.1) else F(500.0))
    difmax = np.full(n, bd, np.float32)
    if name in ("UU", "VV"):
        f = F(1.7) if p < F(401.0) else (F(1.4) if p < F(701.0) else (F(1.1) if p < F(851.0) else None))
        if f is not None:
            difmax[:] = bd * f
    elif name == "PMSLPSFC":
        for k in range(n):
            difmax[k] = bd * (F(1.0) + (F(elev[k]) / F(2000.0)))
    elif name == "TT":
        f = F(0.6) if p < F(401.0) else (F(0.8) if p < F(701.0) else None)
        if f is not None:
            difmax[:] = bd * f
    elif name == "DEWPOINT":
        for k in range(n):
            o = F(obs[k])
            fac = F(1.25) if (o < F(255.0) and o >= F(235.0)) else (F(1.5) if (o < F(235.0) and o >= F(215.0)) else (F(1.75) if o < F(215.0) else F(1.0)))
            difmax[k] = bd * fac
    difmax = (difmax * F(buddy_weight)).astype(np.float32)
    aob = np.array([bilinear(gridded, xob[k], yob[k]) for k in range(n)], np.float32)
    obs = np.asarray(obs, np.float32); xob = np.asarray(xob, np.float32); yob = np.asarray(yob, np.float32)
    qc = np.array(qc_flag, np.int64)
    err = np.zeros(n, np.float32); bnf = np.zeros(n, np.int32)
    dxf = F(dx)
    for k in range(n):
        x = xob[k] - xob; y = yob[k] - yob
        dist = (np.sqrt((x * x + y * y).astype(np.float32)) * dxf).astype(np.float32)
        m = (dist <= rng) & (dist >= F(0.1))
        m[k] = False
        idx = np.nonzero(m)[0]
        diff = (obs[idx] - aob[idx]).astype(np.float32)
        s = F(0.0)
        for d in diff:
            s = s + d
        e = F(0.0)
        if len(idx) >= 2:
            avg = s / F(len(idx))
            kob = len(idx)
            for d, j in zip(diff, idx):
                if d > difmax[j] * F(1.25) and abs(d - avg) > difmax[j]:
                    kob -= 1
                    s = s - d
            bnf[k] = kob
            if kob >= 2:
                e = (obs[k] - aob[k]) - s / F(kob)
            else:
                qc[k] = add_flag(int(qc[k]), NO_BUDDIES)
        else:
            qc[k] = add_flag(int(qc[k]), NO_BUDDIES)
        err[k] = e
        if bnf[k] == 2:
            difmax[k] = F(1.2) * difmax[k]
    for k in range(n):
        if abs(err[k]) > difmax[k]:
            qc[k] = add_flag(int(qc[k]), FAILS_BUDDY)
    return qc.astype(np.int32), aob, err, difmax, bnf


# ---- error_max, ob_density, obs_distance: same discipline (written from the Fortran) -----------
FAILS_ERROR_MAX = 1 << 16


def bilinear(gridded, x, y):
    """qc1_module.F90:197-204 == proc_oa.F90:520-527; gridded (jns, iew) == Fortran gridded(iew,jns)."""
    iob, job = fint(x), fint(y)
    dxob, dyob = F(x) - F(iob), F(y) - F(job)
    g = lambda i, j: F(gridded[j - 1, i - 1])
    one = F(1.0)
    return F(F(one - dxob) * F(F(F(one - dyob) * g(iob, job)) + F(dyob * g(iob, job + 1)))
             + F(dxob * F(F(F(one - dyob) * g(iob + 1, job)) + F(dyob * g(iob + 1, job + 1)))))


def error_max(obs, xob, yob, elev, qc_flag, gridded, name, max_difference, pressure, local_time):
    """src/qc1_module.F90:102-249 (+ modify_qc_flag, qc0_module.F90:9-34).  Returns (qc, aob, err,
    factored_difference)."""
    n = len(obs)
    plevel = F(pressure)
    qc = np.array(qc_flag, np.int32).copy()
    aob = np.zeros(n, np.float32); err = np.zeros(n, np.float32); thr = np.zeros(n, np.float32)
    for k in range(n):
        factor = F(1.0)                                               # :108
        if name in ("UU", "VV"):
            if plevel <= F(499.9):
                factor = F(1.5)
            elif plevel <= F(1000.1):
                factor = F(1.25)
        elif name == "TT":
            if plevel > F(1000.1):
                factor = F(1.5) if 1100 <= int(local_time[k]) <= 2000 else F(1.25)
            elif plevel <= F(1000.1) and plevel > F(700.0):
                factor = F(0.85)
            else:
                factor = F(0.65)
        elif name == "PMSLPSFC":
            factor = F(F(1.0) + F(abs(F(elev[k])) / F(2000.0)))
        elif name == "DEWPOINT":
            o = F(obs[k])
            if o < F(255.0) and o >= F(235.0):
                factor = F(1.25)
            elif o < F(235.0) and o >= F(215.0):
                factor = F(1.50)
            elif o < F(215.0):
                factor = F(1.75)
        aob[k] = bilinear(gridded, xob[k], yob[k])
        err[k] = F(F(obs[k]) - aob[k])                                # :220
        if name in ("UU", "VV"):
            thr[k] = max(F(factor * F(max_difference)), F(F(0.15) * aob[k]))   # :232
        else:
            thr[k] = F(factor * F(max_difference))                    # :241
        if abs(err[k]) > thr[k]:
            qc[k] = add_flag(int(qc[k]), FAILS_ERROR_MAX)
    return qc, aob, err, thr


def ob_density(xob, yob, grid_dist, tobbox):
    """src/qc0_module.F90:81-119; tobbox (iew, jns) C order == Fortran tobbox(jns,iew)."""
    tb = np.array(tobbox, np.float32).copy()
    iew, jns = tb.shape
    r = F(F(250.0) / F(grid_dist))
    rr = nint(r)
    for x, y in zip(xob, yob):
        iob, job = fint(x), fint(y)
        for j in range(max(job - rr - 1, 1), min(job + rr + 1, jns - 1) + 1):
            for i in range(max(iob - rr - 1, 1), min(iob + rr + 1, iew - 1) + 1):
                di, dj = F(F(i) - F(x)), F(F(j) - F(y))
                if np.sqrt(F(F(di * di) + F(dj * dj))) < r:
                    tb[i - 1, j - 1] += F(1.0)
    return tb


def obs_distance(tobbox, dx):
    """src/qc0_module.F90:123-232, the O((iew*jns)**2) search as written."""
    tb = np.asarray(tobbox, np.float32)
    iew, jns = tb.shape
    dx = F(dx)
    a = F(F(F(F(iew - 1) * dx) * F(iew - 1)) * dx); b = F(F(F(F(jns - 1) * dx) * F(jns - 1)) * dx)
    diag_m = np.sqrt(F(a + b))
    diag_cells = F(diag_m / dx)
    obs = np.argwhere(tb > F(0.5))
    out = np.zeros((iew, jns), np.float32)
    for i in range(iew):
        for j in range(jns):
            if tb[i, j] > F(0.5):
                continue
            dijmin = diag_cells
            for ii, jj in obs:
                di, dj = abs(int(ii) - i), abs(int(jj) - j)
                dij = np.sqrt(F(F(F(dj) * F(dj)) + F(F(di) * F(di))))
                dijmin = min(dijmin, dij)
            out[i, j] = F(dijmin * dx)
    return out


def local_time(time_hhmm, lonob):
    """src/qc0_module.F90:259-267 (integer local_time: the REAL expression is truncated on assignment)."""
    out = np.zeros(len(lonob), np.int32)
    for k, lon in enumerate(lonob):
        lt = fint(F(F(time_hhmm // 100) + F(F(lon) / F(15.0))))
        if lt < 0:
            lt = 24 + lt
        out[k] = (lt % 24) * 100
    return out


def contains_2n(sumtocheck, numtofind):
    """src/obs_sort_module.F90:4515-4560: is the power of two `numtofind` a term of the sum of
    distinct powers of two `sumtocheck`?  Greedy decomposition from 2**18 down, as written."""
    curn = 18
    sumnow = int(sumtocheck)
    while sumnow > 0:
        if sumnow < numtofind:
            return False
        if sumnow == numtofind:
            return True
        cur2n = 2 ** curn
        while cur2n > sumnow:
            curn -= 1
            cur2n = 2 ** curn
        if cur2n == numtofind:
            return True
        sumnow -= cur2n
    return False


def qc_consistency(qc_u, qc_v, qc_t, qc_rh):
    """src/qc0_module.F90:371-404 on flat rows (the td > t rule of :405-424 needs the report's
    dew-point measurement and is outside the table)."""
    u, v, t, rh = (np.array(a, np.int32).copy() for a in (qc_u, qc_v, qc_t, qc_rh))
    for k in range(len(u)):
        qu, qv = int(u[k]), int(v[k])
        if qu > qv:
            v[k] = qu
        if qv > qu:
            u[k] = qv
        qt, qr = int(t[k]), int(rh[k])
        contains = contains_2n
        new = qr
        if contains(qt, FAILS_ERROR_MAX) and not contains(qr, FAILS_ERROR_MAX):
            new = add_flag(new, FAILS_ERROR_MAX)
        if contains(qt, FAILS_BUDDY) and not contains(qr, FAILS_BUDDY):
            new = add_flag(new, FAILS_BUDDY)
        rh[k] = new
    return u, v, rh


def scan_radii(radius_influence, sfc_mult, pressure):
    """src/proc_oa.F90:641-707: radii (grid cells, NINT) the multi_scan loop uses for one slab and
    whether the reference STOPs (surface scan 1 below 4.5 cells)."""
    out, stop = [], False
    for num_scan, r in enumerate(radius_influence, 1):
        if r < 0:
            break
        sfc = abs(F(pressure) - F(1001.0)) < F(0.0001)
        sf = F(sfc_mult) if sfc else F(1.00)
        scaled = F(F(r) * sf)
        scan1 = F(F(radius_influence[0]) * sf)
        if F(sfc_mult) < F(0.99999) and sfc:
            if scan1 > F(100.0):
                scaled = F(scaled * F(F(100.0) / scan1))
            if scaled < F(4.5):
                stop = num_scan == 1
                break
        out.append(nint(scaled))
    return out, stop


# ---- smoothers and GET_FG_T_AT_POINT ------------------------------------------------------------
def smooth_5(field, passes, crsdot=1):
    """src/oa_module.F90:928-1006; field (jns, iew) C order == Fortran field(iew,jns)."""
    f = np.array(field, np.float32).copy()
    jns, iew = f.shape
    at = lambda a, i, j: a[j - 1, i - 1]
    for _ in range(passes):
        t = np.zeros_like(f)
        for j in range(2, jns - 1 - crsdot + 1):
            for i in range(2, iew - 1 - crsdot + 1):
                t[j - 1, i - 1] = F(F(F(F(F(F(at(f, i, j) * F(4.0)) + at(f, i + 1, j)) + at(f, i - 1, j)) + at(f, i, j + 1))
                                      + at(f, i, j - 1)) * F(0.125))
        for i in (1, iew - crsdot):
            for j in range(2, jns - 1 - crsdot + 1):
                t[j - 1, i - 1] = F(F(F(F(at(f, i, j) * F(2.0)) + at(f, i, j + 1)) + at(f, i, j - 1)) * F(0.25))
        for j in (1, jns - crsdot):
            for i in range(2, iew - 1 - crsdot + 1):
                t[j - 1, i - 1] = F(F(F(F(at(f, i, j) * F(2.0)) + at(f, i + 1, j)) + at(f, i - 1, j)) * F(0.25))
        for j in range(2, jns - 1 - crsdot + 1):
            for i in range(2, iew - 1 - crsdot + 1):
                f[j - 1, i - 1] = t[j - 1, i - 1]
        for j in range(2, jns - 1 - crsdot + 1):
            f[j - 1, 0] = t[j - 1, 0]; f[j - 1, iew - crsdot - 1] = t[j - 1, iew - crsdot - 1]
        for i in range(2, iew - 1 - crsdot + 1):
            f[0, i - 1] = t[0, i - 1]; f[jns - crsdot - 1, i - 1] = t[jns - crsdot - 1, i - 1]
    return f


def smoother_desmoother(slab, passes, crsdot=1):
    """src/oa_module.F90:875-923; slab (jmx, imx) C order == Fortran slab(imx,jmx)."""
    s = np.array(slab, np.float32).copy()
    jmx, imx = s.shape
    xnu = (F(0.50), F(-0.52))
    half = F(0.5)
    for loop in range(1, passes * 2 + 1):
        nu = xnu[2 - loop % 2 - 1]
        new = s.copy()
        for i in range(2, imx - 1 - crsdot + 1):
            for j in range(2, jmx - 1 - crsdot + 1):
                c = s[j - 1, i - 1]
                new[j - 1, i - 1] = F(c + F(nu * F(F(F(s[j, i - 1] + s[j - 2, i - 1]) * half) - c)))
        s = new
        new = s.copy()
        for j in range(2, jmx - 1 - crsdot + 1):
            for i in range(2, imx - 1 - crsdot + 1):
                c = s[j - 1, i - 1]
                new[j - 1, i - 1] = F(c + F(nu * F(F(F(s[j - 1, i] + s[j - 1, i - 2]) * half) - c)))
        s = new
    return s


def fg_t_at_point(t3d, h3d, x, y, height, dtdz=-0.0065, missing=-888888.0):
    """src/get_fg_at_point_module.F90:36-114; t3d / h3d (kbu, iew, jns) C order == Fortran
    (jns,iew,kbu); (x, y) already projected."""
    kbu, iew, jns = t3d.shape
    x, y, height = F(x), F(y), F(height)
    if x < 1 or x > iew - 1 or y < 1 or y > jns - 1:
        return F(missing)
    fx, cx = int(math.floor(x)), int(math.ceil(x))
    fy, cy = int(math.floor(y)), int(math.ceil(y))
    near = []
    for (ix, iy) in ((fx, fy), (fx, cy), (cx, fy), (cx, cy)):
        hcol = lambda k: F(h3d[k - 1, ix - 1, iy - 1]); tcol = lambda k: F(t3d[k - 1, ix - 1, iy - 1])
        val = None
        for k in range(2, kbu - 1 + 1):
            if height >= hcol(k) and height <= hcol(k + 1):
                w = F(F(hcol(k + 1) - height) / F(hcol(k + 1) - hcol(k)))
                val = F(F(w * tcol(k)) + F(F(F(1.0) - w) * tcol(k + 1)))
                break
        if val is None:
            if height < hcol(2):
                val = F(tcol(2) - F(F(dtdz) * F(hcol(2) - height)))
            elif height > hcol(kbu):
                return F(missing)
            else:
                raise AssertionError("branch the reference STOPs in")
        near.append(val)
    xn, yn = F(x - F(math.floor(x))), F(y - F(math.floor(y)))
    one = F(1.0)
    return F(F(F(one - xn) * F(F(F(one - yn) * near[0]) + F(yn * near[1])))
             + F(xn * F(F(F(one - yn) * near[2]) + F(yn * near[3]))))
