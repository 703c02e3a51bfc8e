// speedy-b200: column physics -- one thread per (grid column, ensemble member), lane = member.
//
// Reference semantics: physics.f90:103-231 driving humidity.f90:44-78, convection.f90:27-253,
// large_scale_condensation.f90:33-96, shortwave_radiation.f90:50-214,325-404, longwave_radiation.f90:16-205,
// surface_fluxes.f90:40-320 and vertical_diffusion.f90:30-146.  The 8-level column lives in registers; the
// per-band transmissivities (rad_tau2) are model state, written on short-wave steps and re-read (L1/L2 hits)
// by the long-wave sweeps, exactly as the reference keeps them in ModelState_t between steps.
// All literal constants carry the reference's REAL(4) values.
#include "kernels.h"

namespace spdy {

namespace ph {
__device__ constexpr double AKAP = (double)(2.0f / 7.0f), CP = FL(1004.0), RGAS = AKAP * CP, P0 = FL(1.e+5),
                            GRAV = FL(9.81), ALHC = FL(2501.0), SBC = FL(5.67e-8), EPSLW = FL(0.05), EMISFC = FL(0.98);
// convection.f90:15-22
__device__ constexpr double PSMIN = FL(0.8), TRCNV = FL(6.0), RHBL = FL(0.9), RHIL = FL(0.7), ENTMAX = FL(0.5), SMF = FL(0.8);
// shortwave_radiation.f90:14-44
__device__ constexpr double RHCL1 = FL(0.30), RHCL2 = FL(1.00), QACL = FL(0.20), WPCL = FL(0.2), PMAXCL = FL(10.0),
                            CLSMAX = FL(0.60), CLSMINL = FL(0.15), GSE_S0 = FL(0.25), GSE_S1 = FL(0.40), ALBCL = FL(0.43),
                            ALBCLS = FL(0.50), ABSDRY = FL(0.033), ABSAER = FL(0.033), ABSWV1 = FL(0.022),
                            ABSWV2 = FL(15.000), ABSCL1 = FL(0.015), ABSCL2 = FL(0.15), ABLWIN = FL(0.3),
                            ABLWV1 = FL(0.7), ABLWV2 = FL(50.0), ABLCL1 = FL(12.0), ABLCL2 = FL(0.6);
// surface_fluxes.f90:13-32
__device__ constexpr double FWIND0 = FL(0.95), FTEMP0 = FL(1.0), CDL = FL(2.4e-3), CDS = FL(1.0e-3), CHL = FL(1.2e-3),
                            CHS = FL(0.9e-3), VGUST = FL(5.0), CTDAY = FL(1.0e-2), DTHETA = FL(3.0), FSTAB = FL(0.67),
                            CLAMBDA = FL(7.0), CLAMBSN = FL(7.0);
}  // namespace ph

// Reciprocal and exponential used by the column physics.  The physics kernel is bound by instruction issue and
// by dependent-FMA latency (2 warps per scheduler), and the libdevice division (~25 instructions with a slow-path
// call) and exp (~40, Horner chain) made up 40 % of its instructions.  Both replacements are accurate to <= 2 ulp
// (validated against the oracle at 1e-12 in tests/test_physics_gpu.py), far inside the parity tolerance.
//   fast_rcp : hardware seed (>= 20 bits) + two Newton steps.
//   fast_exp : k = rint(x log2 e), r = x - k ln2 (two-term Cody-Waite), degree-13 Taylor polynomial evaluated with
//              Estrin's scheme (depth 5 instead of 13), scaled by 2^k through the exponent field.  Results below the
//              normal range are flushed to zero, above it to +inf; NaN propagates.
__device__ __forceinline__ double fast_rcp(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fast_exp(double x) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of (t + MAGIC) is rint(t)
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r, a1 = fma(r, 1.0 / 6.0, 0.5), a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0),
                 a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0), a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0),
                 a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0), a6 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    double y = p * __hiloint2double((k + 1023) << 20, 0);
    y = (x < -708.0) ? 0.0 : y;
    y = (x > 709.0) ? __longlong_as_double(0x7ff0000000000000ll) : y;
    return y;
}
// humidity.f90:44-78 for sig > 0
__device__ __forceinline__ double qsat_of(double ta, double ps, double sig) {
    const double e0 = 6.108e-3, c1 = FL(17.269), c2 = FL(21.875), t0 = FL(273.16), t1 = FL(35.86), t2 = FL(7.66);
    const bool warm = (ta >= t0);
    const double cc = warm ? c1 : c2, tt = warm ? t1 : t2;
    const double q = e0 * fast_exp(cc * (ta - t0) * fast_rcp(ta - tt));
    return FL(622.0) * q * fast_rcp(sig * ps - FL(0.378) * q);
}
// x**3.0 and x**4.0 of longwave_radiation.f90:61,67 and surface_fluxes.f90:216,296: the reference calls the libm
// power function; the products below differ from it by at most 2 ulp (2e-16 relative), far inside the 1e-12 parity
// tolerance, and cost 2 multiplies instead of ~150 instructions each.
__device__ __forceinline__ double pow3(double x) { return x * x * x; }
__device__ __forceinline__ double pow4(double x) { const double x2 = x * x; return x2 * x2; }
// fband(nint(T), jb): the reference indexes fband(100:400,4) unguarded; clamp (documented in DESIGN.md)
__device__ __forceinline__ double fband_at(const double *__restrict__ fb, double T, int jb) {
    long it = lround(T);
    it = it < 100 ? 100 : (it > 400 ? 400 : it);
    return __ldg(fb + (it - 100) + 301 * jb);
}

#ifndef PHYS_MINBLOCKS
#define PHYS_MINBLOCKS 3
#endif
constexpr int PH_SROWS = 7 * KX + 2;  // staged rows per thread: tau2 (32), tt_rsw (8), (accumulated) T tendency (8), q
                                     // tendency (8), lowest-level u and v tendencies (2)
// FUSE: the grid-point dynamics of the same column (grid_dyn_column, dynamics.cu; tendencies.f90:132-224) run first in
// this thread and hand their T / tracer tendencies (and the lowest-level u, v tendencies) over in registers, in the
// order the reference accumulates them (physics.f90 adds to the dynamical tendencies): 34 loads and 34 stores per
// column never reach HBM.
template <bool FUSE>
__global__ void __launch_bounds__(128, PHYS_MINBLOCKS) k_physics(const Ctx c, const ScratchLayout L, int *__restrict__ dbg) {
    using namespace ph;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const int j = q / IX;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
#define ST2D(v) (stp_nc(c, t, c.off[v], lane) + e)
    // Register relief (three CTAs per SM instead of two): the long-wave transmissivities and the short-wave heating
    // (state, read-only on 2 of 3 steps) are copied straight into shared memory with cp.async while the first half of
    // the kernel runs, and the accumulated T tendency waits there between the condensation and the final sum.
    // The dynamical T / tracer tendencies that the physics adds to (needed only after the convection) arrive the same
    // way: first cp.async group, so they cost neither registers nor a stall at the top of the kernel.
    extern __shared__ __align__(16) double ph_sm[];  // PH_SROWS x 128 doubles (58 KB: dynamic, three CTAs per SM)
    double *const sm = ph_sm + threadIdx.x;
    // request the tile index and the lane mask now: every state address below depends on them
    const unsigned mask0 = __ldg(c.masks + t);
    const int tile0 = __ldg(c.tiles + t);
    double *ottend = scp(c, t, L.ttend, lane) + e, *oqtend = scp(c, t, L.trtend, lane) + e;
    if (!FUSE) {
#pragma unroll
        for (int k = 0; k < KX; k++) cp_async8(sm + (5 * KX + k) * 128, ottend + k * lev);
#pragma unroll
        for (int k = 0; k < KX; k++) cp_async8(sm + (6 * KX + k) * 128, oqtend + k * lev);
        cp_async8(sm + (7 * KX) * 128, scp(c, t, L.utend, lane) + e + 7 * lev);
        cp_async8(sm + (7 * KX + 1) * 128, scp(c, t, L.vtend, lane) + e + 7 * lev);
    }
    cp_async_commit();
    // ---- grid-point inputs (physics.f90:89-101): scratch addresses do not depend on the tile list, so these loads are
    // in flight while the first access to the state arena (below) still waits for its tile index
    double ta[KX], qa[KX], phi[KX];
    {
        const double *pt = scp(c, t, L.ptg, lane) + e, *pq = scp(c, t, L.pqg, lane) + e, *pp = scp(c, t, L.pphig, lane) + e;
#pragma unroll
        for (int k = 0; k < KX; k++) ta[k] = pt[k * lev], qa[k] = pq[k * lev], phi[k] = pp[k * lev];
    }
    const double ua8 = *(scp(c, t, L.pug8, lane) + e), va8 = *(scp(c, t, L.pvg8, lane) + e);
    const double psl = *(scp(c, t, L.pslg, lane) + e);
    const bool act = (mask0 >> lane) & 1u;
    const double *fb = c.G->fband;
    const bool do_sw = *(c.st + ((long long)tile0 * c.st_elems + c.off_slots + SL_SW) * TILE + lane) != 0.0;
    if (!(do_sw && act)) {  // on short-wave steps the values are produced below
        const double *pt2 = stp_nc(c, t, c.off[V_rad_tau2], lane) + e, *ptr = stp_nc(c, t, c.off[V_tt_rsw], lane) + e;
#pragma unroll
        for (int k = 0; k < 4 * KX; k++) cp_async8(sm + k * 128, pt2 + k * lev);
#pragma unroll
        for (int k = 0; k < KX; k++) cp_async8(sm + (4 * KX + k) * 128, ptr + k * lev);
    }
    cp_async_commit();
#pragma unroll
    for (int k = 0; k < KX; k++) qa[k] = fmax(qa[k], 0.0);
    const double psa = fast_exp(psl);
    const double rps = fast_rcp(psa);
    double se[KX], rh[KX], qsat[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) {
        se[k] = CP * ta[k] + phi[k];
        qsat[k] = qsat_of(ta[k], psa, c_T.fsg[k]);
        rh[k] = qa[k] * fast_rcp(qsat[k]);
    }
    // T and q tendencies are accumulated in the reference's order of additions and written once
    double tsum[KX], qsum[KX], ut8 = 0.0, vt8 = 0.0;
    if (FUSE) grid_dyn_column<true>(c, L, t, lane, q, tsum, qsum, ut8, vt8);

    // inputs of the later sections (long-wave sweeps, surface fluxes, u/v tendency update): start pulling them into
    // L2 now, so that with only two warps per scheduler those sections wait an L2 hit instead of a DRAM access
    {
        const bool sw = do_sw;
        if (!sw) {  // on short-wave steps these are produced below, not read
            const double *pst = stp_nc(c, t, c.off[V_rad_strat_corr], lane) + e;
            prefetch_l2(pst), prefetch_l2(pst + lev), prefetch_l2(ST2D(V_ssrd));
        } else {
            prefetch_l2(ST2D(V_zenit_correction)), prefetch_l2(ST2D(V_flux_solar_in)), prefetch_l2(ST2D(V_flux_ozone_upper));
            prefetch_l2(ST2D(V_flux_ozone_lower)), prefetch_l2(ST2D(V_alb_surface)), prefetch_l2(ST2D(V_stratospheric_correction));
        }
        prefetch_l2(ST2D(V_phis0)), prefetch_l2(ST2D(V_fmask_land)), prefetch_l2(ST2D(V_forog)), prefetch_l2(ST2D(V_sst_am));
        prefetch_l2(ST2D(V_alb_land)), prefetch_l2(ST2D(V_alb_sea)), prefetch_l2(ST2D(V_snowc));
        prefetch_l2(ST2D(V_land_temp)), prefetch_l2(ST2D(V_soil_avail_water));
    }

    // ---- deep convection (convection.f90:27-253)
    int itop = KX + 1;  // 1-based level index as in the reference, 9 = no convection
    double cbmf = 0.0, precnv = 0.0;
    double dfse[KX], dfqa[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) dfse[k] = 0.0, dfqa[k] = 0.0;
    {
        double qdif = 0.0;
        if (psa > PSMIN) {  // diagnose_convection
            const double mse0 = se[7] + ALHC * qa[7];
            double mse1 = se[6] + ALHC * qa[6];
            mse1 = fmin(mse0, mse1);
            const double mss7 = se[7] + ALHC * qsat[7];
            const double mss0 = fmax(mse0, mss7);
            int ktop1 = KX, ktop2 = KX;
            double msthr = 0.0;
#pragma unroll
            for (int k = KX - 3; k >= 3; k--) {  // 1-based k = 5,4,3
                const double mssk = se[k - 1] + ALHC * qsat[k - 1], mssk1 = se[k] + ALHC * qsat[k];
                const double mss2 = mssk + c_T.wvi[k - 1][1] * (mssk1 - mssk);
                if (mss0 > mss2) ktop1 = k;
                if (mse1 > mss2) ktop2 = k, msthr = mss2;
            }
            if (ktop1 < KX) {
                const double qthr0 = RHBL * qsat[7], qthr1 = RHBL * qsat[6];
                const bool lqthr = (qa[7] > qthr0 && qa[6] > qthr1);
                if (ktop2 < KX) {
                    itop = ktop1;
                    qdif = fmax(qa[7] - qthr0, (mse0 - msthr) * (1.0 / ALHC));
                } else if (lqthr) {
                    itop = ktop1;
                    qdif = qa[7] - qthr0;
                }
            }
        }
        if (itop != KX + 1) {
            // entrainment profile (convection.f90:63-72)
            const double fm0 = c_T.ph_fm0;  // entrainment profile and fm0: host tables (spdy.cuh)
            const double rdps = 2.0 / (1.0 - PSMIN);
            const double qmax = fmax(FL(1.01) * qa[7], qsat[7]);
            double sb = se[6] + c_T.wvi[6][1] * (se[7] - se[6]);
            double qb = qa[6] + c_T.wvi[6][1] * (qa[7] - qa[6]);
            qb = fmin(qb, qa[7]);
            const double fpsa = psa * fmin(1.0, (psa - PSMIN) * rdps);
            double fmass = fm0 * fpsa * fmin(5.0, qdif * fast_rcp(qmax - qb));
            cbmf = fmass;
            double fus = fmass * se[7], fuq = fmass * qmax, fds = fmass * sb, fdq = fmass * qb;
            dfse[7] = fds - fus;
            dfqa[7] = fdq - fuq;
#pragma unroll
            for (int k = KX - 1; k >= 4; k--) {  // 1-based k = 7 .. itop+1 (itop >= 3)
                if (k >= itop + 1) {
                    const int k0 = k - 1, k1 = k - 2;
                    dfse[k0] = fus - fds;
                    dfqa[k0] = fuq - fdq;
                    const double enmass = c_T.ph_entrs[k0] * psa * cbmf;
                    fmass = fmass + enmass;
                    fus = fus + enmass * se[k0];
                    fuq = fuq + enmass * qa[k0];
                    sb = se[k1] + c_T.wvi[k1][1] * (se[k0] - se[k1]);
                    qb = qa[k1] + c_T.wvi[k1][1] * (qa[k0] - qa[k1]);
                    fds = fmass * sb;
                    fdq = fmass * qb;
                    dfse[k0] = dfse[k0] + fds - fus;
                    dfqa[k0] = dfqa[k0] + fdq - fuq;
                    const double delq = RHIL * qsat[k0] - qa[k0];
                    if (delq > 0.0) {
                        const double fsq = SMF * cbmf * delq;
                        dfqa[k0] = dfqa[k0] + fsq;
                        dfqa[7] = dfqa[7] - fsq;
                    }
                }
            }
            // top layer: itop in {3,4,5}
#pragma unroll
            for (int k = 3; k <= 5; k++)
                if (k == itop) {
                    const int k0 = k - 1;
                    const double qsatb = qsat[k0] + c_T.wvi[k0][1] * (qsat[k0 + 1] - qsat[k0]);
                    precnv = fmax(fuq - fmass * qsatb, 0.0);
                    dfse[k0] = fus - fds + ALHC * precnv;
                    dfqa[k0] = fuq - fdq - precnv;
                }
        }
    }
    // physics.f90:128-131,140-141 ; tt_cnv(1) stays 0.  The convective terms wait in registers until the dynamical
    // tendencies (first cp.async group) are needed: the sums keep the reference's order ((dyn + cnv) + lsc) but the
    // wait for the copy moves behind the condensation.
    double tcnv[KX], qcnv[KX];
    tcnv[0] = 0.0, qcnv[0] = 0.0;
#pragma unroll
    for (int k = 1; k < KX; k++) {
        tcnv[k] = dfse[k] * rps * c_T.grdscp[k];
        qcnv[k] = dfqa[k] * rps * c_T.grdsig[k];
    }
    const int icnv = KX - itop;

    // ---- large-scale condensation (large_scale_condensation.f90:33-96)
    double precls = 0.0;
    {
        const double rtlsc = 1.0 / (FL(4.0) * FL(3600.0));
        const double tfact = ALHC / CP, prg = P0 / GRAV;
        const double psa2 = psa * psa;
        double dtl[KX], dql[KX];
        dtl[0] = 0.0, dql[0] = 0.0;
#pragma unroll
        for (int k = 1; k < KX; k++) {
            const double sig2 = c_T.fsg[k] * c_T.fsg[k];
            double rhref = FL(0.9) + FL(0.1) * (sig2 - 1.0);
            if (k == KX - 1) rhref = fmax(rhref, FL(0.95));
            const double dqmax = 10.0 * sig2 * rtlsc;
            const double dqa = rhref * qsat[k] - qa[k];
            if (dqa < 0.0) {
                itop = min(k + 1, itop);
                dql[k] = dqa * rtlsc;
                dtl[k] = tfact * fmin(-dql[k], dqmax * psa2);
            } else {
                dql[k] = 0.0, dtl[k] = 0.0;
            }
        }
#pragma unroll
        for (int k = 1; k < KX; k++) precls = precls - (c_T.dhs[k] * prg) * dql[k];
        precls = precls * psa;
        // physics.f90:140-141: ttend = ttend + tt_cnv + tt_lsc
        if (!FUSE) {
            cp_async_wait<1>();  // first group: the dynamical tendencies (each thread reads only what it copied itself)
#pragma unroll
            for (int k = 0; k < KX; k++) tsum[k] = sm[(5 * KX + k) * 128], qsum[k] = sm[(6 * KX + k) * 128];
        }
#pragma unroll
        for (int k = 0; k < KX; k++) tsum[k] = (tsum[k] + tcnv[k]) + dtl[k], qsum[k] = (qsum[k] + qcnv[k]) + dql[k];
#pragma unroll
        for (int k = 0; k < KX; k++) sm[(5 * KX + k) * 128] = tsum[k];
    }
    const bool dg = act && c.diag_out;  // outputs nothing reads before the next step overwrites them (Ctx::diag_out)
    if (dg) {
        *ST2D(V_cbmf) = cbmf;
        *ST2D(V_precnv) = precnv;
        *ST2D(V_precls) = precls;
    }

    // ---- radiation
    double *tau2 = stp_nc(c, t, c.off[V_rad_tau2], lane) + e;  // (ix,il,kx,4): element (k, b) at (k + KX*b)*lev
    double *ttrsw = stp_nc(c, t, c.off[V_tt_rsw], lane) + e;
    double *strat = stp_nc(c, t, c.off[V_rad_strat_corr], lane) + e;
    int icltop_out = 0;
    if (do_sw && act) {  // physics.f90:151-169 ; inactive lanes own no state
        // clouds (shortwave_radiation.f90:325-404)
        const double gse = (se[6] - se[7]) * fast_rcp(phi[6] - phi[7]);
        double cloudc;
        int icltop;
        if (rh[6] > RHCL1) cloudc = rh[6] - RHCL1, icltop = KX - 1;
        else cloudc = 0.0, icltop = KX + 1;
#pragma unroll
        for (int k = 3; k <= KX - 2; k++) {
            const double drh = rh[k - 1] - RHCL1;
            if (drh > cloudc && qa[k - 1] > QACL) cloudc = drh, icltop = k;
        }
        {
            const double pr1 = fmin(PMAXCL, FL(86.4) * (precnv + precls));
            const double c2 = fmin(1.0, cloudc * (1.0 / (RHCL2 - RHCL1)));
            cloudc = fmin(1.0, WPCL * sqrt(pr1) + c2 * c2);
            icltop = min(itop, icltop);
        }
        icltop_out = icltop;
        const double qcloud = qa[6];
        const double fmask = *ST2D(V_fmask_land);
        double clstr;
        {
            const double fstab = fmax(0.0, fmin(1.0, (1.0 / (GSE_S1 - GSE_S0)) * (gse - GSE_S0)));
            clstr = fstab * fmax(CLSMAX - FL(1.2) * cloudc, 0.0);
            const double clstrl = fmax(clstr, CLSMINL) * rh[7];
            clstr = clstr + fmask * (clstrl - clstr);
        }
        *ST2D(V_qcloud_equiv) = qcloud;
        // get_shortwave_rad_fluxes (shortwave_radiation.f90:50-214)
        const double fband2 = FL(0.05), fband1 = 1.0 - fband2;
        const double psaz = psa * *ST2D(V_zenit_correction);
        const double acloud = cloudc * fmin(ABSCL1 * qcloud, ABSCL2);
        const double fsol = *ST2D(V_flux_solar_in);
        double tau1[KX], tau3[KX], trsw[KX];
        tau1[0] = fast_exp(-psaz * c_T.dhs[0] * ABSDRY);
#pragma unroll
        for (int k = 1; k < KX - 1; k++) {
            const double abs1 = ABSDRY + ABSAER * (c_T.fsg[k] * c_T.fsg[k]);
            if (k + 1 >= icltop) tau1[k] = fast_exp(-psaz * c_T.dhs[k] * (abs1 + ABSWV1 * qa[k] + acloud));
            else tau1[k] = fast_exp(-psaz * c_T.dhs[k] * (abs1 + ABSWV1 * qa[k]));
        }
        {
            const double abs1 = ABSDRY + ABSAER * (c_T.fsg[7] * c_T.fsg[7]);
            tau1[7] = fast_exp(-psaz * c_T.dhs[7] * (abs1 + ABSWV1 * qa[7]));
        }
#pragma unroll
        for (int k = 0; k < KX; k++) tau3[k] = 0.0;
#pragma unroll
        for (int k = 1; k <= KX; k++)
            if (k == icltop) tau3[k - 1] = ALBCL * cloudc;
        tau3[7] = ALBCLS * clstr;
        double tsr = fsol, f1 = fsol * fband1, f2 = fsol * fband2;
        // stratosphere
        trsw[0] = f1;
        f1 = tau1[0] * (f1 - *ST2D(V_flux_ozone_upper) * psa);
        trsw[0] = trsw[0] - f1;
        trsw[1] = f1;
        f1 = tau1[1] * (f1 - *ST2D(V_flux_ozone_lower) * psa);
        trsw[1] = trsw[1] - f1;
        // troposphere
#pragma unroll
        for (int k = 2; k < KX; k++) {
            tau3[k] = f1 * tau3[k];
            f1 = f1 - tau3[k];
            trsw[k] = f1;
            f1 = tau1[k] * f1;
            trsw[k] = trsw[k] - f1;
        }
#pragma unroll
        for (int k = 1; k < KX; k++) {
            const double tau2k = fast_exp(-psaz * c_T.dhs[k] * ABSWV2 * qa[k]);
            trsw[k] = trsw[k] + f2;
            f2 = tau2k * f2;
            trsw[k] = trsw[k] - f2;
        }
        const double ssrd = f1 + f2;
        f1 = f1 * *ST2D(V_alb_surface);
        const double ssr = ssrd - f1;
#pragma unroll
        for (int k = KX - 1; k >= 0; k--) {
            trsw[k] = trsw[k] + f1;
            f1 = tau1[k] * f1;
            trsw[k] = trsw[k] - f1;
            f1 = f1 + tau3[k];
        }
        tsr = tsr - f1;
        *ST2D(V_tsr) = tsr;
        *ST2D(V_ssrd) = ssrd;
        *ST2D(V_ssr) = ssr;
        // the reference leaves the SW-phase rad_flux(:,:,1:2) in state until the LW sweep overwrites them
        // long-wave transmissivities (section 5)
        const double co2 = *stp_nc(c, t, c.off_slots + SL_CO2, lane);
        const double acl2 = cloudc * ABLCL2;
#pragma unroll
        for (int k = 0; k < KX; k++) {
            double t1, t2, t3, t4;
            const double deltap = psa * c_T.dhs[k];
            if (k == 0) {
                t1 = fast_exp(-psa * c_T.dhs[k] * ABLWIN), t2 = fast_exp(-psa * c_T.dhs[k] * co2), t3 = 1.0, t4 = 1.0;
            } else if (k == 1 || k == KX - 1) {
                t1 = fast_exp(-psa * c_T.dhs[k] * ABLWIN), t2 = fast_exp(-psa * c_T.dhs[k] * co2);
                t3 = fast_exp(-psa * c_T.dhs[k] * ABLWV1 * qa[k]), t4 = fast_exp(-psa * c_T.dhs[k] * ABLWV2 * qa[k]);
            } else {
                const double acloud1 = (k + 1 < icltop) ? acl2 : ABLCL1 * cloudc;
                t1 = fast_exp(-deltap * (ABLWIN + acloud1)), t2 = fast_exp(-deltap * co2);
                t3 = fast_exp(-deltap * fmax(ABLWV1 * qa[k], acl2)), t4 = fast_exp(-deltap * fmax(ABLWV2 * qa[k], acl2));
            }
            tau2[(k + KX * 0) * lev] = t1, tau2[(k + KX * 1) * lev] = t2;
            tau2[(k + KX * 2) * lev] = t3, tau2[(k + KX * 3) * lev] = t4;
            const double trk = trsw[k] * rps * c_T.grdscp[k];  // physics.f90:166-168
            ttrsw[k * lev] = trk;
            sm[(k + KX * 0) * 128] = t1, sm[(k + KX * 1) * 128] = t2, sm[(k + KX * 2) * 128] = t3, sm[(k + KX * 3) * 128] = t4;
            sm[(4 * KX + k) * 128] = trk;
        }
        const double eps1 = c_T.ph_eps1;
        strat[0] = *ST2D(V_stratospheric_correction) * psa;
        strat[lev] = eps1 * psa;
    }

    // ---- vertical diffusion and shallow convection (vertical_diffusion.f90:30-146): independent of the radiation,
    //      evaluated here so that se/rh/qsat/phi die before the long-wave sweeps (register pressure)
    double tv[KX], qv[KX];
    {
        const double redshc = FL(0.5), segrad = FL(0.1);
        const double fshcq = c_T.ph_fshcq, fshcse = c_T.ph_fshcse, fvdise = c_T.ph_fvdise;  // host tables
#pragma unroll
        for (int k = 0; k < KX; k++) tv[k] = 0.0, qv[k] = 0.0;
        const double rsig6 = c_T.ph_rdhs[6], rsig7 = c_T.ph_rdhs[7];
        {
            const double drh0 = c_T.ph_drh0[7];
            const double fvdiq2 = c_T.ph_fvdiq2[7];
            const double dmse = se[7] - se[6] + ALHC * (qa[7] - qsat[6]);
            const double drh = rh[7] - rh[6];
            double fcnv = 1.0;
            if (dmse >= 0.0) {
                if (icnv > 0) fcnv = redshc;
                const double fluxse = fcnv * fshcse * dmse;
                tv[6] = fluxse * rsig6;
                tv[7] = -fluxse * rsig7;
                if (drh >= 0.0) {
                    const double fluxq = fcnv * fshcq * qsat[7] * drh;
                    qv[6] = fluxq * rsig6;
                    qv[7] = -fluxq * rsig7;
                }
            } else if (drh > drh0) {
                const double fluxq = fvdiq2 * qsat[6] * drh;
                qv[6] = fluxq * rsig6;
                qv[7] = -fluxq * rsig7;
            }
        }
#pragma unroll
        for (int k = 3; k <= KX - 2; k++)  // 1-based k
            if (c_T.sigh[k] > 0.5) {
                const double drh0 = c_T.ph_drh0[k];
                const double fvdiq2 = c_T.ph_fvdiq2[k];
                const double drh = rh[k] - rh[k - 1];
                if (drh >= drh0) {
                    const double fluxq = fvdiq2 * qsat[k - 1] * drh;
                    qv[k - 1] = qv[k - 1] + fluxq * c_T.ph_rdhs[k - 1];
                    qv[k] = qv[k] - fluxq * c_T.ph_rdhs[k];
                }
            }
#pragma unroll
        for (int k = 0; k < KX - 1; k++) {
            const double se0 = se[k + 1] + segrad * (phi[k] - phi[k + 1]);
            if (se[k] < se0) {
                const double fluxse = fvdise * (se0 - se[k]);
                tv[k] = tv[k] + fluxse * c_T.ph_rdhs[k];
                const double r1 = c_T.ph_r1sig[k];
#pragma unroll
                for (int k1 = k + 1; k1 < KX; k1++) tv[k1] = tv[k1] - fluxse * r1;
            }
        }
    }
    // moisture tendency: final for levels 1..kx-1 (the lowest level still needs the evaporation)
#pragma unroll
    for (int k = 0; k < KX - 1; k++) oqtend[k * lev] = qsum[k] + qv[k];
    const double ta6 = ta[6], ta7 = ta[7], qa7 = qa[7], phi7 = phi[7];

    // ---- downward long-wave (longwave_radiation.f90:16-121)
    double st4a1[KX], st4a2[KX], dfabs[KX], flux[4];
    cp_async_wait<0>();  // staged transmissivities (each thread reads only what it copied itself)
#define TAU(jb, k) sm[((k) + KX * (jb)) * 128]
    {
#pragma unroll
        for (int k = 0; k < KX - 1; k++) st4a1[k] = ta[k] + c_T.wvi[k][1] * (ta[k + 1] - ta[k]);
        st4a2[0] = FL(0.75) * ta[0] + FL(0.25) * st4a1[0];
        st4a2[1] = FL(0.50) * ta[1] + FL(0.25) * (st4a1[0] + st4a1[1]);
#pragma unroll
        for (int k = 2; k < KX - 1; k++) st4a2[k] = 0.5 * 1.0 * fmax(st4a1[k] - st4a1[k - 1], 0.0);
        st4a2[KX - 1] = 1.0 * fmax(ta[KX - 1] - st4a1[KX - 2], 0.0);
#pragma unroll
        for (int k = 0; k < 2; k++) st4a1[k] = SBC * pow4(st4a2[k]), st4a2[k] = 0.0;
#pragma unroll
        for (int k = 2; k < KX; k++) {
            const double st3a = SBC * pow3(ta[k]);
            st4a1[k] = st3a * ta[k];
            st4a2[k] = 4.0 * st3a * st4a2[k];
        }
#pragma unroll
        for (int k = 0; k < KX; k++) dfabs[k] = 0.0;
#pragma unroll
        for (int jb = 0; jb < 2; jb++) {
            const double emis = 1.0 - TAU(jb, 0);
            const double brad = fband_at(fb, ta[0], jb) * (st4a1[0] + emis * st4a2[0]);
            flux[jb] = emis * brad;
            dfabs[0] = dfabs[0] - flux[jb];
        }
        flux[2] = 0.0, flux[3] = 0.0;
#pragma unroll
        for (int jb = 0; jb < 4; jb++)
#pragma unroll
            for (int k = 1; k < KX; k++) {
                const double tk = TAU(jb, k);
                const double emis = 1.0 - tk;
                const double brad = fband_at(fb, ta[k], jb) * (st4a1[k] + emis * st4a2[k]);
                dfabs[k] = dfabs[k] + flux[jb];
                flux[jb] = tk * flux[jb] + emis * brad;
                dfabs[k] = dfabs[k] - flux[jb];
            }
    }
    double slrd = 0.0;
#pragma unroll
    for (int jb = 0; jb < 4; jb++) slrd = slrd + EMISFC * flux[jb];
    {
        const double corlw = EPSLW * EMISFC * st4a1[KX - 1];
        dfabs[KX - 1] = dfabs[KX - 1] - corlw;
        slrd = slrd + corlw;
    }

    // ---- surface fluxes (surface_fluxes.f90:40-320, lfluxland = .true.)
    double ts, shf3, evap3, ustr3, vstr3, slru3;
    {
        const double phi0 = *ST2D(V_phis0), fmask = *ST2D(V_fmask_land), forog = *ST2D(V_forog),
                     tsea = *ST2D(V_sst_am), ssrd = *ST2D(V_ssrd), alb_land = *ST2D(V_alb_land),
                     alb_sea = *ST2D(V_alb_sea), snowc = *ST2D(V_snowc), land_temp = *ST2D(V_land_temp),
                     saw = *ST2D(V_soil_avail_water);
        const double esbc = EMISFC * SBC;
        const double u0 = FWIND0 * ua8, v0 = FWIND0 * va8;
        const double gtemp0 = 1.0 - FTEMP0, rcp = 1.0 / CP;
        const double dt1 = c_T.wvi[KX - 1][1] * (ta7 - ta6);
        double t1l = ta7 + dt1;
        double t1s = t1l - phi0 * dt1 * c_T.ph_rt1s;
        const double t2s = ta7 + rcp * phi7;
        const double t2l = t2s - rcp * phi0;
        if (ta7 > ta6) {
            t1l = FTEMP0 * t1l + gtemp0 * t2l;
            t1s = FTEMP0 * t1s + gtemp0 * t2s;
        } else {
            t1l = ta7, t1s = ta7;
        }
        const double t0 = t1s + fmask * (t1l - t1s);
        const double denvvs0 = (P0 * psa * fast_rcp(RGAS * t0)) * sqrt(u0 * u0 + v0 * v0 + VGUST * VGUST);
        double tskin = land_temp + CTDAY * c_T.ph_sqcoa[j] * ssrd * (1.0 - alb_land) * psa;
        const double rdth = FSTAB / DTHETA, astab = 0.5;
        const double dthl = (tskin > t2l) ? fmin(DTHETA, tskin - t2l) : fmax(-DTHETA, astab * (tskin - t2l));
        const double denvvs1 = denvvs0 * (1.0 + dthl * rdth);
        const double cdldv = CDL * denvvs0 * forog;
        const double ustr1 = -cdldv * ua8, vstr1 = -cdldv * va8;
        const double chlcp = CHL * CP;
        double shf1 = chlcp * denvvs1 * (tskin - t1l);
        const double q1 = qa7;
        const double qsat01 = qsat_of(tskin, psa, 1.0);
        double evap1 = CHL * denvvs1 * fmax(0.0, saw * qsat01 - q1);
        const double tsk3 = pow3(tskin);
        const double dslr = 4.0 * esbc * tsk3;
        double slru1 = esbc * tsk3 * tskin;
        double hfl1 = ssrd * (1.0 - alb_land) + slrd - (slru1 + shf1 + ALHC * evap1);
        const double clamb = CLAMBDA + snowc * (CLAMBSN - CLAMBDA);
        hfl1 = hfl1 - clamb * (tskin - land_temp);
        double qsat02 = qsat_of(tskin + 1.0, psa, 1.0);
        qsat02 = (evap1 > 0.0) ? saw * (qsat02 - qsat01) : 0.0;
        const double dtskin = hfl1 * fast_rcp(clamb + dslr + CHL * denvvs1 * (CP + ALHC * qsat02));
        tskin = tskin + dtskin;
        shf1 = shf1 + chlcp * denvvs1 * dtskin;
        evap1 = evap1 + CHL * denvvs1 * qsat02 * dtskin;
        slru1 = slru1 + dslr * dtskin;
        hfl1 = clamb * (tskin - land_temp);
        const double dths = (tsea > t2s) ? fmin(DTHETA, tsea - t2s) : fmax(-DTHETA, astab * (tsea - t2s));
        const double denvvs2 = denvvs0 * (1.0 + dths * rdth);
        const double cdsdv = CDS * denvvs2;
        const double ustr2 = -cdsdv * ua8, vstr2 = -cdsdv * va8;
        const double shf2 = CHS * CP * denvvs2 * (tsea - t1s);
        const double qsats = qsat_of(tsea, psa, 1.0);
        const double evap2 = CHS * denvvs2 * (qsats - q1);
        const double slru2 = esbc * pow4(tsea);
        const double hfl2 = ssrd * (1.0 - alb_sea) + slrd - slru2 + shf2 + ALHC * evap2;
        ustr3 = ustr2 + fmask * (ustr1 - ustr2);
        vstr3 = vstr2 + fmask * (vstr1 - vstr2);
        shf3 = shf2 + fmask * (shf1 - shf2);
        evap3 = evap2 + fmask * (evap1 - evap2);
        slru3 = slru2 + fmask * (slru1 - slru2);
        ts = tsea + fmask * (land_temp - tsea);
        if (act) {  // the land / sea models (surface.cu) read hfluxn, shf(sea), evap(sea)
            double *p;
            p = ST2D(V_shf), p[lev] = shf2;
            p = ST2D(V_evap), p[lev] = evap2;
            p = ST2D(V_hfluxn), p[0] = hfl1, p[lev] = hfl2;
        }
        if (dg) {
            double *p;
            p = ST2D(V_ustr), p[0] = ustr1, p[lev] = ustr2, p[2 * lev] = ustr3;
            p = ST2D(V_vstr), p[0] = vstr1, p[lev] = vstr2, p[2 * lev] = vstr3;
            p = ST2D(V_shf), p[0] = shf1, p[2 * lev] = shf3;
            p = ST2D(V_evap), p[0] = evap1, p[2 * lev] = evap3;
            p = ST2D(V_slru), p[0] = slru1, p[lev] = slru2, p[2 * lev] = slru3;
            *ST2D(V_slrd) = slrd;
        }
    }

    // ---- upward long-wave (longwave_radiation.f90:124-205)
    {
        const double refsfc = 1.0 - EMISFC;
        const double fsfcu = slru3;
        const double slr = fsfcu - slrd;
#pragma unroll
        for (int jb = 0; jb < 4; jb++) flux[jb] = fband_at(fb, ts, jb) * fsfcu + refsfc * flux[jb];
        dfabs[KX - 1] = dfabs[KX - 1] + EPSLW * fsfcu;
#pragma unroll
        for (int jb = 0; jb < 4; jb++)
#pragma unroll
            for (int k = KX - 1; k >= 1; k--) {
                const double tk = TAU(jb, k);
                const double emis = 1.0 - tk;
                const double brad = fband_at(fb, ta[k], jb) * (st4a1[k] - emis * st4a2[k]);
                dfabs[k] = dfabs[k] + flux[jb];
                flux[jb] = tk * flux[jb] + emis * brad;
                dfabs[k] = dfabs[k] - flux[jb];
            }
#pragma unroll
        for (int jb = 0; jb < 2; jb++) {
            const double tk = TAU(jb, 0);
            const double emis = 1.0 - tk;
            const double brad = fband_at(fb, ta[0], jb) * (st4a1[0] - emis * st4a2[0]);
            dfabs[0] = dfabs[0] + flux[jb];
            flux[jb] = tk * flux[jb] + emis * brad;
            dfabs[0] = dfabs[0] - flux[jb];
        }
        const double s1 = strat[0], s2 = strat[lev];
        const double corlw1 = c_T.dhs[0] * s2 * st4a1[0] + s1;
        const double corlw2 = c_T.dhs[1] * s2 * st4a1[1];
        dfabs[0] = dfabs[0] - corlw1;
        dfabs[1] = dfabs[1] - corlw2;
        double olr = corlw1 + corlw2;
#pragma unroll
        for (int jb = 0; jb < 4; jb++) olr = olr + flux[jb];
        if (dg) {
            *ST2D(V_slr) = slr;
            *ST2D(V_olr) = olr;
            double *pf = ST2D(V_rad_flux), *ps4 = ST2D(V_rad_st4a);
#pragma unroll
            for (int jb = 0; jb < 4; jb++) pf[jb * lev] = flux[jb];
#pragma unroll
            for (int k = 0; k < KX; k++) ps4[k * lev] = st4a1[k], ps4[(k + KX) * lev] = st4a2[k];
        }
    }
    // physics.f90:200-204: ttend = ttend + tt_rsw + tt_rlw ; :214-231: + surface fluxes and PBL tendencies
    {
        const double utp = 0.0 + ustr3 * rps * c_T.grdsig[7], vtp = 0.0 + vstr3 * rps * c_T.grdsig[7];
        tv[7] = tv[7] + shf3 * rps * c_T.grdscp[7];
        qv[7] = qv[7] + evap3 * rps * c_T.grdsig[7];
        double *outend = scp(c, t, L.utend, lane) + e + 7 * lev, *ovtend = scp(c, t, L.vtend, lane) + e + 7 * lev;
        *outend = (FUSE ? ut8 : sm[(7 * KX) * 128]) + utp;
        *ovtend = (FUSE ? vt8 : sm[(7 * KX + 1) * 128]) + vtp;
        oqtend[7 * lev] = qsum[7] + qv[7];
#pragma unroll
        for (int k = 0; k < KX; k++)
            ottend[k * lev] = ((sm[(5 * KX + k) * 128] + sm[(4 * KX + k) * 128]) + dfabs[k] * rps * c_T.grdscp[k]) + tv[k];
    }

    if (dbg) {
        int *d = dbg + ((size_t)t * 3 * NG + q) * TILE + lane;
        d[0] = itop, d[(size_t)NG * TILE] = icnv, d[(size_t)2 * NG * TILE] = icltop_out;
    }
#undef ST2D
#undef TAU
}

constexpr int PH_SMEM = PH_SROWS * 128 * 8;
template <bool FUSE> static void launch_physics_t(cudaStream_t s, const Ctx &c, const ScratchLayout &L, int *dbg) {
    static bool init = false;
    if (!init) {
        if (cudaFuncSetAttribute(k_physics<FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PH_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_physics needs %d bytes of shared memory per CTA (sm_100a)\n", PH_SMEM);
            abort();
        }
        init = true;
    }
    k_physics<FUSE><<<dim3(NG / 4, c.ntiles), 128, PH_SMEM, s>>>(c, L, dbg);
}
void launch_dyn_physics(cudaStream_t s, const Ctx &c, const ScratchLayout &L) { launch_physics_t<true>(s, c, L, nullptr); }
void launch_physics(cudaStream_t s, const Ctx &c, const ScratchLayout &L, int *dbg) { launch_physics_t<false>(s, c, L, dbg); }

}  // namespace spdy
