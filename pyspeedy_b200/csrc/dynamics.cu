// speedy-b200: grid-point dynamics, spectral tendencies + semi-implicit + diffusion + leapfrog, diagnostics,
// per-member control (calendar) kernels.
//
// Reference semantics: tendencies.f90:51-352, implicit.f90:234-289, horizontal_diffusion.f90:131-152,
// time_stepping.f90:38-188, diagnostics.f90:16-74, model_control.f90:113-186, speedy.f90:20-74.
// One thread = one (grid column | spectral coefficient) of one member; lane = member.
#include "kernels.h"

namespace spdy {

// physical constants with the reference's REAL(4)-literal values (physical_constants.f90:15-47)
__device__ constexpr double D_AKAP = (double)(2.0f / 7.0f);
__device__ constexpr double D_CP = FL(1004.0);
__device__ constexpr double D_RGAS = D_AKAP * D_CP;
__device__ constexpr double D_ROB = FL(0.05), D_WIL = FL(0.53);

// ------------------------------------------------------------------------------------ grid-point dynamics
// tendencies.f90:132-224.  Inputs: ug, vg, tg, vorg, divg, trg (level j2), px, py.  Outputs: utend, vtend,
// ttend, trtend and the grid field whose transform is psdt.
// FUSE: the column is handed on to the physics of the same thread (k_physics<true>, physics.cu): the T and tracer
// tendencies and the lowest-level u, v tendencies stay in registers instead of making a round trip through HBM.
template <bool FUSE>
__device__ __forceinline__ void grid_dyn_column(const Ctx &c, const ScratchLayout &L, const int t, const int lane, const int q,
                                                double (&tt)[KX], double (&qt)[KX], double &ut8, double &vt8) {
    const int j = q / IX;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    const double *pu = scp(c, t, L.ug, lane) + e, *pv = scp(c, t, L.vg, lane) + e, *pt = scp(c, t, L.tg, lane) + e,
                 *pvo = scp(c, t, L.vorg, lane) + e, *pd = scp(c, t, L.divg, lane) + e, *pq = scp(c, t, L.trg, lane) + e;
    double u[KX], v[KX], T[KX], vo[KX], d[KX], tr[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) {
        u[k] = pu[k * lev], v[k] = pv[k * lev], T[k] = pt[k * lev];
        vo[k] = pvo[k * lev] + c_T.coriol[j];  // :126-130
        d[k] = pd[k * lev], tr[k] = pq[k * lev];
    }
    const double px = *(scp(c, t, L.px, lane) + e), py = *(scp(c, t, L.py, lane) + e);
    double umean = 0.0, vmean = 0.0, dmean = 0.0;
#pragma unroll
    for (int k = 0; k < KX; k++) {
        umean = umean + u[k] * c_T.dhs[k];
        vmean = vmean + v[k] * c_T.dhs[k];
        dmean = dmean + d[k] * c_T.dhs[k];
    }
    *(scp(c, t, L.psdtg, lane) + e) = -umean * px - vmean * py;  // :148
    double puv[KX], sigdt[KX + 1], sigm[KX + 1], tgg[KX];
    sigdt[0] = 0.0, sigm[0] = 0.0;
#pragma unroll
    for (int k = 0; k < KX; k++) {
        puv[k] = (u[k] - umean) * px + (v[k] - vmean) * py;
        sigdt[k + 1] = sigdt[k] - c_T.dhs[k] * (puv[k] + d[k] - dmean);
        sigm[k + 1] = sigm[k] - c_T.dhs[k] * puv[k];
        tgg[k] = T[k] - c_T.tref[k];
    }
    double *ou = scp(c, t, L.utend, lane) + e, *ov = scp(c, t, L.vtend, lane) + e, *ot = scp(c, t, L.ttend, lane) + e,
           *oq = scp(c, t, L.trtend, lane) + e;
    double tmp[KX + 1];
    tmp[0] = 0.0, tmp[KX] = 0.0;
    // zonal wind :174-184
#pragma unroll
    for (int k = 1; k < KX; k++) tmp[k] = sigdt[k] * (u[k] - u[k - 1]);
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double r = v[k] * vo[k] - tgg[k] * D_RGAS * px - (tmp[k + 1] + tmp[k]) * c_T.dhsr[k];
        if (FUSE && k == KX - 1) ut8 = r;
        else ou[k * lev] = r;
    }
    // meridional wind :187-194
#pragma unroll
    for (int k = 1; k < KX; k++) tmp[k] = sigdt[k] * (v[k] - v[k - 1]);
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double r = -u[k] * vo[k] - tgg[k] * D_RGAS * py - (tmp[k + 1] + tmp[k]) * c_T.dhsr[k];
        if (FUSE && k == KX - 1) vt8 = r;
        else ov[k * lev] = r;
    }
    // temperature :197-209
#pragma unroll
    for (int k = 1; k < KX; k++) tmp[k] = sigdt[k] * (tgg[k] - tgg[k - 1]) + sigm[k] * (c_T.tref[k] - c_T.tref[k - 1]);
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double r = tgg[k] * d[k] - (tmp[k + 1] + tmp[k]) * c_T.dhsr[k] + c_T.fsgr[k] * tgg[k] * (sigdt[k + 1] + sigdt[k]) +
                         c_T.tref3[k] * (sigm[k + 1] + sigm[k]) + D_AKAP * (T[k] * puv[k] - tgg[k] * dmean);
        if (FUSE) tt[k] = r;
        else ot[k * lev] = r;
    }
    // tracer :212-224 (temp(:,:,2:3) = 0)
#pragma unroll
    for (int k = 1; k < KX; k++) tmp[k] = (k <= 2) ? 0.0 : sigdt[k] * (tr[k] - tr[k - 1]);
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double r = tr[k] * d[k] - (tmp[k + 1] + tmp[k]) * c_T.dhsr[k];
        if (FUSE) qt[k] = r;
        else oq[k * lev] = r;
    }
}
__global__ void __launch_bounds__(128) k_grid_dyn(const Ctx c, const ScratchLayout L) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    double tt[KX], qt[KX], ut8, vt8;
    grid_dyn_column<false>(c, L, t, lane, q, tt, qt, ut8, vt8);
}

// -------------------------------------------------------------------------- spectral tendencies + time step
// Complex helpers (used by the grid<->spectral service kernels)
struct C2 {
    double r, i;
};
__device__ __forceinline__ C2 ld2(const double *p) { return {p[0], p[TILE]}; }
__device__ __forceinline__ void st2(double *p, C2 v) { p[0] = v.r, p[TILE] = v.i; }
__device__ __forceinline__ C2 operator+(C2 a, C2 b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ C2 operator-(C2 a, C2 b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ C2 operator*(double s, C2 a) { return {s * a.r, s * a.i}; }
__device__ __forceinline__ C2 operator*(C2 a, double s) { return {a.r * s, a.i * s}; }
__device__ __forceinline__ C2 operator-(C2 a) { return {-a.r, -a.i}; }

// (vor, div) of vel2vort at coefficient (m,n) from spectral u, v (spectral.f90:160-186), complex form
__device__ __forceinline__ void vdspec_elem(const GlobTables *G, const double *su, const double *sv, int m, int n,
                                            C2 &vor, C2 &dv) {
    const int q = m + MX * n;
    const size_t up = (size_t)M2 * TILE;
    const double gx = G->gradx[m];
    const C2 uc = ld2(su), vc = ld2(sv);
    const C2 zp = {-(gx * uc.i), gx * uc.r}, zc = {-(gx * vc.i), gx * vc.r};
    if (n == 0) {
        const double yp = G->vddyp[q];
        vor = zc - yp * ld2(su + up);
        dv = zp + yp * ld2(sv + up);
    } else if (n == NX - 1) {
        const double ym = G->vddym[q];
        vor = ym * ld2(su - up);
        dv = (-ym) * ld2(sv - up);
    } else {
        const double ym = G->vddym[q], yp = G->vddyp[q];
        vor = (ym * ld2(su - up) - yp * ld2(su + up)) + zc;
        dv = ((-ym) * ld2(sv - up) + yp * ld2(sv + up)) + zp;
    }
}

// Same operator for ONE component (cc = 0 real, 1 imaginary), branch-free so that the loads of all levels can be
// issued together.  su/sv point at the (m,n) element of this component; the i*gradx term needs the other component of
// the centre element only (`oth` = +TILE / -TILE).  The special rows n = 1 and n = nx of spectral.f90:171-176 are
// the general row with exact zeros: vddym(m,1) = 0 and vddyp(m,nx) = 0 in the tables (spectral.f90:98-99,110 with
// epsi(:,nx+1) = 0), the neighbour that does not exist is replaced by the element itself (multiplied by that zero),
// and row nx has no i*gradx term (zsel = 0).  WANT: 1 = vorticity, 2 = divergence.
template <int WANT>
__device__ __forceinline__ void vdspec_comp(const double gx, const double ym, const double yp, const double *su,
                                            const double *sv, const int n, const int cc, double &vor, double &dv) {
    const long long up = (n == NX - 1) ? 0 : (long long)M2 * TILE, dn = (n == 0) ? 0 : -(long long)M2 * TILE;
    const long long oth = cc ? -TILE : TILE;
    const double zsel = (n == NX - 1) ? 0.0 : (cc ? gx : -gx);  // (a+bi)*i = -b + ai
    if (WANT & 1) vor = (ym * su[dn] - yp * su[up]) + zsel * sv[oth];
    if (WANT & 2) dv = ((-ym) * sv[dn] + yp * sv[up]) + zsel * su[oth];
}

// leapfrog + Robert-Asselin-Williams filter on one component (time_stepping.f90:164-188)
// (o1, o2 = current values of the two time levels, loaded by the caller so that loads can be batched)
__device__ __forceinline__ void raw_update(double *p1, double *p2, double o1, double o2, double fdt, const double trf,
                                           const int j1, const double dt, const double eps, const bool act) {
    fdt = fdt * trf;  // truncate (ix == 4*iy)
    const double fnew = o1 + dt * fdt;
    double oj1 = (j1 == 1) ? o1 : o2;
    o1 = oj1 + (D_WIL * eps) * ((o1 - 2.0 * oj1) + fnew);
    if (j1 == 1) oj1 = o1;  // output(:,:,j1) aliases the updated level 1
    o2 = fnew - ((1.0 - D_WIL) * eps) * ((o1 - 2.0 * oj1) + fnew);
    if (act) *p1 = o1, *p2 = o2;
}

// raw_update with a zero tendency for the coefficients outside the truncation (see k_spec_step_vq).  The two time levels
// are written only where the filter changed them: they are exact zeros there unless the host stored something else, and
// zeros stay zeros, so these rows cost their reads only.
__device__ __forceinline__ void raw_update_idle(double *p1, double *p2, const double a1, const double a2, const double trf,
                                                const int j1, const double dt, const double eps, const bool act) {
    double o1 = a1, o2 = a2;
    const double fdt = 0.0 * trf;
    const double fnew = o1 + dt * fdt;
    double oj1 = (j1 == 1) ? o1 : o2;
    o1 = oj1 + (D_WIL * eps) * ((o1 - 2.0 * oj1) + fnew);
    if (j1 == 1) oj1 = o1;
    o2 = fnew - ((1.0 - D_WIL) * eps) * ((o1 - 2.0 * oj1) + fnew);
    if (act && __double_as_longlong(o1) != __double_as_longlong(a1)) *p1 = o1;
    if (act && __double_as_longlong(o2) != __double_as_longlong(a2)) *p2 = o2;
}

// Outside the triangular truncation the tendency is zero (trfilt), and the Robert-Asselin-Williams filter with a zero
// tendency is the identity on a coefficient whose two time levels hold the same bits: (a - 2a) + a == 0 exactly, so both
// levels keep a.  (Exceptions: -0, which the filter turns into +0, and NaN.)  That is the normal state of those rows -- zeros,
// and in row m + n = 31 whatever grid2spectral left there at initialisation -- unless the host stored something else.  A
// multi-step driver call therefore looks at them ONCE (here) and its spectral steps skip those rows for the tiles where
// every active member passed (Ctx::outer_zero).  Two flags per tile, preset to non-zero: [2t] row m + n = 31 (a perturbed
// ensemble set up like examples/Ensemble_forecast.ipynb -- grid2spectral writes time level 1 only -- keeps moving there),
// [2t + 1] the 465 coefficients with m + n >= 32.
__device__ __forceinline__ bool outer_moves(const double a1, const double a2) {
    const long long b1 = __double_as_longlong(a1), b2 = __double_as_longlong(a2);
    return b1 != b2 || b1 == (long long)0x8000000000000000ull || a1 != a1;
}
__global__ void __launch_bounds__(256) k_scan_outer(const Ctx c, int *__restrict__ flags) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 8 + (threadIdx.x >> 5), t = blockIdx.y;
    const int cc = w & 1, q = w >> 1;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    if (m + n <= NTRUNC) return;
    const size_t e = (size_t)(2 * m + cc + M2 * n) * TILE, lev = (size_t)NSP * TILE, tl = (size_t)KX * lev;
    const double *vor = stp(c, t, c.off[V_vor], lane) + e, *dvs = stp(c, t, c.off[V_div], lane) + e,
                 *tt = stp(c, t, c.off[V_t], lane) + e, *trs = stp(c, t, c.off[V_tr], lane) + e,
                 *ps = stp(c, t, c.off[V_ps], lane) + e;
    bool bad = outer_moves(ps[0], ps[lev]);
#pragma unroll
    for (int k = 0; k < KX; k++)
        bad |= outer_moves(vor[k * lev], vor[tl + k * lev]) | outer_moves(dvs[k * lev], dvs[tl + k * lev]) |
               outer_moves(tt[k * lev], tt[tl + k * lev]) | outer_moves(trs[k * lev], trs[tl + k * lev]);
    if (lane_active(c, t, lane) && bad) flags[2 * t + (m + n > NTRUNC + 1)] = 0;
}

// Vorticity and tracer: no vertical coupling -> one thread per (coefficient, component, level)
// (tendencies.f90:238-268 spectral part, time_stepping.f90:78-144)
__global__ void __launch_bounds__(256) k_spec_step_vq(const Ctx c, const ScratchLayout L, const int j1, const double dt,
                                                      const double eps, const int impl_idx, const long long dump) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 8 + (threadIdx.x >> 5), t = blockIdx.y;
    // w = ((q * 2 + cc) * KX + k)
    const int k = w % KX, cc = (w / KX) & 1, q = w / (2 * KX);
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    const GlobTables *G = c.G;
    const ImplTables *I = &G->impl[impl_idx];
    const size_t e = (size_t)(2 * m + cc + M2 * n) * TILE, lev = (size_t)NSP * TILE, tl = (size_t)KX * lev;
    const bool act = lane_active(c, t, lane);
    const double *F = scp(c, t, L.sfwd, lane) + e;
    double *vor = stp(c, t, c.off[V_vor], lane) + e + k * lev, *trs = stp(c, t, c.off[V_tr], lane) + e + k * lev;
    const double trf = G->trfilt[q];
    // Outside the triangular truncation (m + n > 30, half of the (31,32) array) step_field multiplies the whole tendency
    // by trfilt = 0 (time_stepping.f90:177-179, spectral.f90:72-83): those coefficients only go through the time filter
    // with a zero tendency, so their 8 tendency rows, the correction field and the damping tables are never loaded.
    const bool live = (m + n <= NTRUNC) || dump >= 0;
    if (!live) {
        if (c.outer_zero && c.outer_zero[2 * t + (m + n > NTRUNC + 1)]) return;  // the filter is the identity here (k_scan_outer)
        const double v1 = *vor, q1 = *trs, v2 = vor[tl], q2 = trs[tl];
        raw_update_idle(vor, vor + tl, v1, v2, trf, j1, dt, eps, act);
        raw_update_idle(trs, trs + tl, q1, q2, trf, j1, dt, eps, act);
        return;
    }
    const double gx = G->gradx[m], ym = G->vddym[q], yp = G->vddyp[q];
    double vo, dq, dum;
    vdspec_comp<1>(gx, ym, yp, F + (FW_SU + k) * lev, F + (FW_SV + k) * lev, n, cc, vo, dum);
    vdspec_comp<2>(gx, ym, yp, F + (FW_UQ + k) * lev, F + (FW_VQ + k) * lev, n, cc, dum, dq);
    double trdt = dq + F[(FW_QT + k) * lev];
    if (dump >= 0) {  // test hook: tendencies as returned by get_tendencies (tendencies.f90:11-39), state untouched
        *(scp(c, t, dump + (long long)(0 + k) * NSP, lane) + e) = vo;
        *(scp(c, t, dump + (long long)(25 + k) * NSP, lane) + e) = trdt;
        return;
    }
    const double v1 = *vor, q1 = *trs, v2 = vor[tl], q2 = trs[tl];
    const double dmp = G->dmp[q], dmpd = G->dmpd[q], dmps = G->dmps[q];
    double vordt = (vo - dmp * v1) * I->dmp1[q];
    if (k == 0 && m == 0) vordt = vordt - (1.0 / ((double)(24.0f * 30.0f) * FL(3600.0))) * v1;
    vordt = (vordt - dmps * v1) * I->dmp1s[q];
    const double ctq = q1 + *(stp(c, t, c.off_qcorh, lane) + e) * c_T.qcorv[k];
    trdt = (trdt - dmpd * ctq) * I->dmp1d[q];
    raw_update(vor, vor + tl, v1, v2, vordt, trf, j1, dt, eps, act);
    raw_update(trs, trs + tl, q1, q2, trdt, trf, j1, dt, eps, act);
}

// Divergence, temperature, log(ps): coupled in the vertical by the semi-implicit scheme -> one thread per
// (coefficient, component) holding the 8-level columns (tendencies.f90:283-352, implicit.f90:234-289)
template <bool CT>  // CT: implicit matrices from __constant__ memory (regular step); else from global (first_step)
__global__ void __launch_bounds__(128) k_spec_step_dt(const Ctx c, const ScratchLayout L, const int j1, const double dt,
                                                      const double eps, const int impl_idx, const long long dump) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const int cc = w & 1, q = w >> 1;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    const GlobTables *G = c.G;
    const ImplTables *I = &G->impl[impl_idx];
    const size_t e = (size_t)(2 * m + cc + M2 * n) * TILE, lev = (size_t)NSP * TILE, tl = (size_t)KX * lev;
    const bool act = lane_active(c, t, lane);
    const double *F = scp(c, t, L.sfwd, lane) + e;
    double *dvs = stp(c, t, c.off[V_div], lane) + e, *tt = stp(c, t, c.off[V_t], lane) + e,
           *ps = stp(c, t, c.off[V_ps], lane) + e;
    const double *phi = stp(c, t, c.off[V_phi], lane) + e;
    const double el2 = G->el2[q], trf = G->trfilt[q];
    if (m + n > NTRUNC && dump < 0) {  // outside the truncation: zero tendency, time filter only (see k_spec_step_vq)
        if (c.outer_zero && c.outer_zero[2 * t + (m + n > NTRUNC + 1)]) return;
        double a1[2 * KX + 1], a2[2 * KX + 1];
#pragma unroll
        for (int k = 0; k < KX; k++) {
            a1[k] = dvs[k * lev], a2[k] = dvs[tl + k * lev];
            a1[KX + k] = tt[k * lev], a2[KX + k] = tt[tl + k * lev];
        }
        a1[2 * KX] = ps[0], a2[2 * KX] = ps[lev];
        raw_update_idle(ps, ps + lev, a1[2 * KX], a2[2 * KX], trf, j1, dt, eps, act);
#pragma unroll
        for (int k = 0; k < KX; k++) {
            raw_update_idle(dvs + k * lev, dvs + tl + k * lev, a1[k], a2[k], trf, j1, dt, eps, act);
            raw_update_idle(tt + k * lev, tt + tl + k * lev, a1[KX + k], a2[KX + k], trf, j1, dt, eps, act);
        }
        return;
    }
    const double gx = G->gradx[m], ym = G->vddym[q], yp = G->vddyp[q];
    // rows needed only after the (long) implicit solve: start pulling them into L2 now
    if (dump < 0) {
#pragma unroll
        for (int k = 0; k < KX; k++) {
            prefetch_l2(phi + k * lev), prefetch_l2(tt + k * lev);
            prefetch_l2(dvs + tl + k * lev), prefetch_l2(tt + tl + k * lev);
        }
        prefetch_l2(ps), prefetch_l2(ps + lev), prefetch_l2(stp(c, t, c.off_tcorh, lane) + e);
    }

    double divdt[KX], tdt[KX], d1[KX];
    // ---- A. grid-point tendencies in spectral space (tendencies.f90:238-268)
#pragma unroll
    for (int k = 0; k < KX; k++) {
        double dv, dT, dum;
        vdspec_comp<2>(gx, ym, yp, F + (FW_SU + k) * lev, F + (FW_SV + k) * lev, n, cc, dum, dv);
        divdt[k] = dv - ((-F[(FW_KE + k) * lev]) * el2);
        vdspec_comp<2>(gx, ym, yp, F + (FW_UT + k) * lev, F + (FW_VT + k) * lev, n, cc, dum, dT);
        tdt[k] = dT + F[(FW_TT + k) * lev];
        d1[k] = dvs[k * lev];
    }
    double psdt = F[FW_PS * lev];
    if (q == 0) psdt = 0.0;
    // ---- B. spectral tendencies with time level 1 (tendencies.f90:283-352)
    double dmeanc = 0.0;
#pragma unroll
    for (int k = 0; k < KX; k++) dmeanc = dmeanc + d1[k] * c_T.dhs[k];
    psdt = psdt - dmeanc;
    if (q == 0) psdt = 0.0;
    {
        double sig[KX + 1], dumk[KX + 1];
        sig[0] = 0.0, sig[KX] = 0.0;
#pragma unroll
        for (int k = 0; k < KX - 1; k++) sig[k + 1] = sig[k] - c_T.dhs[k] * (d1[k] - dmeanc);
        dumk[0] = 0.0, dumk[KX] = 0.0;
#pragma unroll
        for (int k = 1; k < KX; k++) dumk[k] = sig[k] * (c_T.tref[k] - c_T.tref[k - 1]);
#pragma unroll
        for (int k = 0; k < KX; k++)
            tdt[k] = ((tdt[k] - (dumk[k + 1] + dumk[k]) * c_T.dhsr[k]) + c_T.tref3[k] * (sig[k + 1] + sig[k])) -
                     c_T.tref2[k] * dmeanc;
    }
    const double ps1 = *ps;
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double g = phi[k * lev] + (D_RGAS * c_T.tref[k]) * ps1;
        divdt[k] = divdt[k] - ((-g) * el2);
    }
    if (dump >= 0 && j1 < 0) {  // test hook, stage 1 (j1 = -1): divdt, tdt, psdt BEFORE the implicit correction, i.e. the
#pragma unroll              // operands implicit_terms receives (time_stepping.f90:71-75)
        for (int k = 0; k < KX; k++) {
            *(scp(c, t, dump + (long long)(8 + k) * NSP, lane) + e) = divdt[k];
            *(scp(c, t, dump + (long long)(16 + k) * NSP, lane) + e) = tdt[k];
        }
        *(scp(c, t, dump + 24ll * NSP, lane) + e) = psdt;
        return;
    }
    // ---- C. semi-implicit correction (implicit.f90:234-289)
    {
        double yf[KX];
        const double elz = I->elz[q];
#pragma unroll
        for (int k = 0; k < KX; k++) {
            double ye = 0.0;
#pragma unroll
            for (int k1 = 0; k1 < KX; k1++) ye = ye + (CT ? c_T.xd2[k + KX * k1] : I->xd[k + KX * k1]) * tdt[k1];
            ye = ye + (D_RGAS * c_T.tref[k]) * psdt;
            yf[k] = divdt[k] + elz * ye;
        }
        const int l = m + n;
#pragma unroll
        for (int k = 0; k < KX; k++) divdt[k] = 0.0;
        if (l != 0) {
            const double *xj = (CT ? c_T.xj2 : I->xj) + (size_t)KX * KX * (l - 1);
#pragma unroll
            for (int k1 = 0; k1 < KX; k1++)
#pragma unroll
                for (int k = 0; k < KX; k++) divdt[k] = divdt[k] + xj[k + KX * k1] * yf[k1];
        }
#pragma unroll
        for (int k = 0; k < KX; k++) psdt = psdt - divdt[k] * (CT ? c_T.dhsx2[k] : I->dhsx[k]);
#pragma unroll
        for (int k = 0; k < KX; k++)
#pragma unroll
            for (int k1 = 0; k1 < KX; k1++) tdt[k] = tdt[k] + (CT ? c_T.xc2[k + KX * k1] : I->xc[k + KX * k1]) * divdt[k1];
    }
    if (dump >= 0) {  // test hook (see k_spec_step_vq): divdt, tdt, psdt after the implicit correction
#pragma unroll
        for (int k = 0; k < KX; k++) {
            *(scp(c, t, dump + (long long)(8 + k) * NSP, lane) + e) = divdt[k];
            *(scp(c, t, dump + (long long)(16 + k) * NSP, lane) + e) = tdt[k];
        }
        *(scp(c, t, dump + 24ll * NSP, lane) + e) = psdt;
        return;
    }
    // ---- D/E. horizontal diffusion, stratospheric drag and time integration (time_stepping.f90:78-144)
    const double dmp = G->dmp[q], dmpd = G->dmpd[q], dmps = G->dmps[q];
    const double dmp1 = I->dmp1[q], dmp1d = I->dmp1d[q], dmp1s = I->dmp1s[q];
    const double sdrag = 1.0 / ((double)(24.0f * 30.0f) * FL(3600.0));
    const double tcorh = *(stp(c, t, c.off_tcorh, lane) + e);
    // all read-modify-write operands are loaded before the first store (loads behind a store would serialise)
    double t1[KX], t2[KX], d2[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) t1[k] = tt[k * lev], t2[k] = tt[tl + k * lev], d2[k] = dvs[tl + k * lev];
    const double ps2 = ps[lev];
    raw_update(ps, ps + lev, ps1, ps2, psdt, trf, j1, dt, eps, act);
#pragma unroll
    for (int k = 0; k < KX; k++) {
        double dd = (divdt[k] - dmpd * d1[k]) * dmp1d;
        const double ctmp = t1[k] + tcorh * c_T.tcorv[k];
        double td = (tdt[k] - dmp * ctmp) * dmp1;
        if (k == 0 && m == 0) dd = dd - sdrag * d1[k];
        dd = (dd - dmps * d1[k]) * dmp1s;
        td = (td - dmps * ctmp) * dmp1s;
        raw_update(dvs + k * lev, dvs + tl + k * lev, d1[k], d2[k], dd, trf, j1, dt, eps, act);
        raw_update(tt + k * lev, tt + tl + k * lev, t1[k], t2[k], td, trf, j1, dt, eps, act);
    }
}

// ------------------------------------------------------------------------------------------------ diagnostics
// diagnostics.f90:16-74 in two deterministic stages: (1) one warp per (level, m) sums over n; (2) one warp per tile
// adds the 30 partial sums in m order and applies the range check.
__global__ void __launch_bounds__(32) k_diag_partial(const Ctx c, const int time_lev, const long long part) {
    const int lane = threadIdx.x, k = blockIdx.x / (MX - 1), m = 1 + blockIdx.x % (MX - 1), t = blockIdx.y;
    const size_t lev = (size_t)NSP * TILE, tl = (size_t)KX * lev * (time_lev - 1);
    const double *vor = stp(c, t, c.off[V_vor], lane) + tl + k * lev, *dv = stp(c, t, c.off[V_div], lane) + tl + k * lev;
    const GlobTables *G = c.G;
    double d1 = 0.0, d2 = 0.0;
#pragma unroll 8
    for (int n = 0; n < NX; n++) {
        const size_t e = (size_t)(2 * m + M2 * n) * TILE;
        const double em = G->elm2[m + MX * n];
        const double vr = vor[e], vi = vor[e + TILE], dr = dv[e], di = dv[e + TILE];
        const double ar = (-vr) * em, ai = (-vi) * em, br = (-dr) * em, bi = (-di) * em;
        d1 = d1 - (ar * vr - ai * (-vi));
        d2 = d2 - (br * dr - bi * (-di));
    }
    double *p = scp(c, t, part + 2 * (k * (MX - 1) + (m - 1)), lane);
    p[0] = d1, p[TILE] = d2;
}
// Second stage + the member bookkeeping that follows the check in do_single_step (speedy.f90:59-69).  One CTA per
// tile, warp k sums level k: its 30 partial pairs are loaded together and added in the serial order m = 1..30.
// mode 0: check only (initialisation).  mode 1: step counter + 1 before the check, calendar advance after it for
// the members that passed (model_control.f90:113-163).
__device__ __forceinline__ void advance_calendar(const Ctx &c, int t, int lane);
// err_out / masks (mode 1): sticky error codes of the driver call and the call's active-lane masks -- a member that fails is
// taken out of the rest of a multi-step call (see step_members in engine.cu)
__global__ void __launch_bounds__(256) k_diag_final(const Ctx c, const int time_lev, const long long part, const int mode,
                                                    int *__restrict__ err_out, unsigned *__restrict__ masks) {
    __shared__ int s_bad[KX][TILE];
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5, t = blockIdx.x;
    const size_t lev = (size_t)NSP * TILE, tl = (size_t)KX * lev * (time_lev - 1);
    const double *tt = stp(c, t, c.off[V_t], lane) + tl;
    double p1[MX - 1], p2[MX - 1];
#pragma unroll
    for (int m = 0; m < MX - 1; m++) {
        const double *p = scp(c, t, part + 2 * (k * (MX - 1) + m), lane);
        p1[m] = p[0], p2[m] = p[TILE];
    }
    const double t00 = tt[k * lev];
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int m = 0; m < MX - 1; m++) d1 += p1[m], d2 += p2[m];
    const double d3 = 0x1.6a09e6p-1 * t00;  // sqrt(0.5) REAL(4)
    s_bad[k][lane] = (d1 > 500.0 || d2 > 500.0 || d3 < 180.0 || d3 > 320.0) ? 1 : 0;
    __syncthreads();
    if (k != 0 || !lane_active(c, t, lane)) return;
    int bad = 0;
#pragma unroll
    for (int kk = 0; kk < KX; kk++) bad |= s_bad[kk][lane];
    if (mode == 1) slot(c, t, lane, SL_STEP) = slot(c, t, lane, SL_STEP) + 1.0;
    if (bad) slot(c, t, lane, SL_ERR) = -2.0;
    const int code = (int)slot(c, t, lane, SL_ERR);
    if (mode == 1 && code == 0) advance_calendar(c, t, lane);
    if (err_out && code != 0 && err_out[t * TILE + lane] == 0) {
        err_out[t * TILE + lane] = code;
        atomicAnd(&masks[t], ~(1u << lane));
    }
}

// --------------------------------------------------------------------------------------------- member control
// speedy.f90:41-53: per-member flags for this step
__device__ __forceinline__ double date_code(const Ctx &c, int t, int lane) {
    return (slot(c, t, lane, SL_YEAR) * 16.0 + slot(c, t, lane, SL_MONTH)) * 32.0 + slot(c, t, lane, SL_DAY);
}
__device__ __forceinline__ void control_pre_lane(const Ctx &c, const int t, const int lane) {
    if (!lane_active(c, t, lane)) return;
    const int step = (int)slot(c, t, lane, SL_STEP);
    // coupler climatology cache: valid for the current date unless the host touched the member since the last step
    slot(c, t, lane, SL_CPLSTAMP) = (slot(c, t, lane, SL_CPLDIRTY) != 0.0) ? -1.0 : date_code(c, t, lane);
    slot(c, t, lane, SL_CPLDIRTY) = 0.0;
    slot(c, t, lane, SL_ERR) = 0.0;
    slot(c, t, lane, SL_DAILY) = (step % NSTEPS == 0) ? 1.0 : 0.0;
    slot(c, t, lane, SL_SW) = (step % NSTRAD == 0) ? 1.0 : 0.0;
}
__global__ void k_control_pre(const Ctx c) { control_pre_lane(c, blockIdx.x, threadIdx.x); }

// The spectral pre-operators of a step in ONE launch, warp per coefficient (m,n): set_geopotential of time level 1
// (geopotential.f90:36-77), vort2vel of time level j2 for the dynamics and of the lowest level of time level 1 for the
// physics (spectral.f90:190-214), grad(ps) (spectral.f90:275-296); with_control: the first warp of a tile also sets the
// member flags of the step (speedy.f90:41-53; on daily steps a separate launch does it before the forcing kernels).
// tri: only the rows n <= 31 - m that the fused inverse transform reads.
__global__ void __launch_bounds__(128) k_preops(const Ctx c, const ScratchLayout L, const int j2, const int tri,
                                                const int with_control) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (with_control && q == 0) control_pre_lane(c, t, lane);
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE, lev = (size_t)NSP * TILE;
    geopotential_elem(stp(c, t, c.off[V_t], lane) + e, stp(c, t, c.off[V_phis], lane) + e, stp(c, t, c.off[V_phi], lane) + e, m,
                      lane_active(c, t, lane));
    if (tri && m + n > NTRUNC + 1) return;
    const size_t tl2 = (size_t)(j2 - 1) * KX * lev;
    const double *pv = stp(c, t, c.off[V_vor], lane), *pd = stp(c, t, c.off[V_div], lane);
    double *pu = scp(c, t, L.ucos, lane), *pw = scp(c, t, L.vcos, lane);
#pragma unroll 2
    for (int k = 0; k < KX; k++) uvspec_elem(c.G, pv + tl2 + k * lev, pd + tl2 + k * lev, pu + k * lev, pw + k * lev, m, n);
    uvspec_elem(c.G, pv + 7 * lev, pd + 7 * lev, scp(c, t, L.ucosp8, lane), scp(c, t, L.vcosp8, lane), m, n);
    gradient_elem(c.G, stp(c, t, c.off[V_ps], lane) + (size_t)(j2 - 1) * lev, scp(c, t, L.dpx, lane), scp(c, t, L.dpy, lane), m, n);
}

__device__ __forceinline__ int days_in_month(int mth) {
    const int d[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    return d[mth - 1];
}
// model_control.f90:166-186 (REAL(4) arithmetic)
__device__ __forceinline__ void update_forcing_params(const Ctx &c, int t, int lane) {
    const int month = (int)slot(c, t, lane, SL_MONTH), day = (int)slot(c, t, lane, SL_DAY);
    int cum = 0;
    for (int mm = 1; mm < month; mm++) cum += days_in_month(mm);
    slot(c, t, lane, SL_IMONT1) = month;
    slot(c, t, lane, SL_TMONTH) = (double)(((float)day - 0.5f) / (float)days_in_month(month));
    slot(c, t, lane, SL_TYEAR) = (double)(((float)(cum + day) - 0.5f) / 365.0f);
}
// speedy.f90:59-69 + model_control.f90:113-163: step counter, then (if the check passed) the calendar
__device__ __forceinline__ void advance_calendar(const Ctx &c, int t, int lane) {
    int year = (int)slot(c, t, lane, SL_YEAR), month = (int)slot(c, t, lane, SL_MONTH),
        day = (int)slot(c, t, lane, SL_DAY), hour = (int)slot(c, t, lane, SL_HOUR),
        minute = (int)slot(c, t, lane, SL_MINUTE), midx = (int)slot(c, t, lane, SL_MONTH_IDX);
    minute += 24 * 60 / NSTEPS;
    if (minute >= 60) minute %= 60, hour += 1;
    if (hour >= 24) hour %= 24, day += 1;
    if (year % 4 == 0 && month == 2) {
        if (day > 29) day = 1, month += 1, midx += 1;
    } else if (day > days_in_month(month)) {
        day = 1, month += 1, midx += 1;
    }
    if (month > 12) month = 1, year += 1;
    slot(c, t, lane, SL_YEAR) = year, slot(c, t, lane, SL_MONTH) = month, slot(c, t, lane, SL_DAY) = day;
    slot(c, t, lane, SL_HOUR) = hour, slot(c, t, lane, SL_MINUTE) = minute, slot(c, t, lane, SL_MONTH_IDX) = midx;
    update_forcing_params(c, t, lane);
}
__global__ void k_update_forcing_params(const Ctx c) {
    const int lane = threadIdx.x, t = blockIdx.x;
    if (lane_active(c, t, lane)) update_forcing_params(c, t, lane);
}

// ---- stand-alone spectral operators for the operator-level entry points (spdy_batch_vel2vort, spdy_batch_laplacian) ----
// vel2vort (spectral.f90:160-186) through the very device function the spectral step uses (vdspec_comp), one component
// per thread
__global__ void __launch_bounds__(128) k_vdspec(const Ctx c, FieldRef u, FieldRef v, FieldRef vor, FieldRef dv) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const int cc = w & 1, q = w >> 1;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    const size_t e = (size_t)(2 * m + cc + M2 * n) * TILE;
    double vo, d;
    vdspec_comp<3>(c.G->gradx[m], c.G->vddym[q], c.G->vddyp[q], refp(c, t, u, lane) + e, refp(c, t, v, lane) + e, n, cc, vo, d);
    *(refp(c, t, vor, lane) + e) = vo, *(refp(c, t, dv, lane) + e) = d;
}
// laplacian / laplacian_inv (spectral.f90:140-155) in the form the step evaluates them: -(x) * el2, -(x) * elm2
__global__ void __launch_bounds__(128) k_laplacian(const Ctx c, FieldRef in, FieldRef out, int inverse) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const int cc = w & 1, q = w >> 1;
    if (q >= NSPC) return;
    const size_t e = (size_t)(2 * (q % MX) + cc + M2 * (q / MX)) * TILE;
    *(refp(c, t, out, lane) + e) = (-*(refp(c, t, in, lane) + e)) * (inverse ? c.G->elm2[q] : c.G->el2[q]);
}
void launch_vdspec(cudaStream_t s, const Ctx &c, FieldRef u, FieldRef v, FieldRef vor, FieldRef dv) {
    k_vdspec<<<dim3(NSPC * 2 / 4, c.ntiles), 128, 0, s>>>(c, u, v, vor, dv);
}
void launch_laplacian(cudaStream_t s, const Ctx &c, FieldRef in, FieldRef out, int inverse) {
    k_laplacian<<<dim3(NSPC * 2 / 4, c.ntiles), 128, 0, s>>>(c, in, out, inverse);
}

void launch_grid_dyn(cudaStream_t s, const Ctx &c, const ScratchLayout &L) {
    k_grid_dyn<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, L);
}
void launch_spec_step(cudaStream_t s, const Ctx &c, const ScratchLayout &L, int j1, double dt, double eps, int impl_idx,
                      long long dump = -1) {
    k_spec_step_vq<<<dim3(NSPC * 2 * KX / 8, c.ntiles), 256, 0, s>>>(c, L, j1, dt, eps, impl_idx, dump);
    if (impl_idx == 2) k_spec_step_dt<true><<<dim3(NSPC * 2 / 4, c.ntiles), 128, 0, s>>>(c, L, j1, dt, eps, impl_idx, dump);
    else k_spec_step_dt<false><<<dim3(NSPC * 2 / 4, c.ntiles), 128, 0, s>>>(c, L, j1, dt, eps, impl_idx, dump);
}
void launch_scan_outer(cudaStream_t s, const Ctx &c, int *flags) {
    k_scan_outer<<<dim3(NSPC * 2 / 8, c.ntiles), 256, 0, s>>>(c, flags);
}
void launch_diag(cudaStream_t s, const Ctx &c, int time_lev, long long part, int mode, int *err_out = nullptr,
                 unsigned *masks = nullptr) {
    k_diag_partial<<<dim3(KX * (MX - 1), c.ntiles), 32, 0, s>>>(c, time_lev, part);
    k_diag_final<<<c.ntiles, 256, 0, s>>>(c, time_lev, part, mode, err_out, masks);
}
void launch_preops(cudaStream_t s, const Ctx &c, const ScratchLayout &L, int j2, int tri, int with_control) {
    k_preops<<<dim3(NSPC / 4, c.ntiles), 128, 0, s>>>(c, L, j2, tri, with_control);
}
void launch_control_pre(cudaStream_t s, const Ctx &c) { k_control_pre<<<c.ntiles, 32, 0, s>>>(c); }
void launch_update_forcing_params(cudaStream_t s, const Ctx &c) { k_update_forcing_params<<<c.ntiles, 32, 0, s>>>(c); }

}  // namespace spdy
