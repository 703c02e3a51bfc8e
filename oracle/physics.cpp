// ORACLE (test infrastructure) -- grid-point column physics.
// Follows humidity.f90, convection.f90, large_scale_condensation.f90, shortwave_radiation.f90,
// longwave_radiation.f90, surface_fluxes.f90, vertical_diffusion.f90 and physics.f90 of the reference.
#include <algorithm>

#include "speedy_oracle.hpp"

namespace orc {

static const size_t NG = (size_t)ix * il;
static inline double dmin(double a, double b) { return a < b ? a : b; }   // Fortran MIN(a,b)
static inline double dmax(double a, double b) { return a > b ? a : b; }   // Fortran MAX(a,b)
static inline long nint(double x) { return lround(x); }                  // Fortran NINT: half away from zero

// ---------------------------------------------------------------------------------------------------
// humidity.f90:44-78.  ps is read element-wise when sig > 0, ps[0] = ps(1,1) otherwise.
void get_qsat(const double *ta, const double *ps, double sig, double *qsat, int n) {
    const double e0 = 6.108e-3, c1 = FL(17.269), c2 = FL(21.875), t0 = FL(273.16), t1 = FL(35.86), t2 = FL(7.66);
    for (int q = 0; q < n; q++) {
        if (ta[q] >= t0)
            qsat[q] = e0 * exp(c1 * (ta[q] - t0) / (ta[q] - t1));
        else
            qsat[q] = e0 * exp(c2 * (ta[q] - t0) / (ta[q] - t2));
    }
    if (sig <= 0.0) {
        double ps11 = ps[0];
        for (int q = 0; q < n; q++) qsat[q] = FL(622.0) * qsat[q] / (ps11 - FL(0.378) * qsat[q]);
    } else {
        for (int q = 0; q < n; q++) qsat[q] = FL(622.0) * qsat[q] / (sig * ps[q] - FL(0.378) * qsat[q]);
    }
}

// ---------------------------------------------------------------------------------------------------
// convection.f90:15-22
static const double psmin = FL(0.8), trcnv = FL(6.0), rhbl = FL(0.9), rhil = FL(0.7), entmax = FL(0.5), smf = FL(0.8);

// convection.f90:170-253
static void diagnose_convection(G2 psa, G3 se, G3 qa, G3 qsat, int *itop, G2 qdif, const Geometry &g) {
    Grid3 mss;  // levels 2..kx used
    double msthr = 0.0;
    const int nl1 = kx - 1, nlp = kx + 1;
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) mss.d[q + NG * (k - 1)] = se.p[q + NG * (k - 1)] + alhc * qsat.p[q + NG * (k - 1)];
    const double rlhc = 1.0 / alhc;
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            int &it = itop[(i - 1) + ix * (j - 1)];
            it = nlp;
            if (psa(i, j) > psmin) {
                double mse0 = se(i, j, kx) + alhc * qa(i, j, kx);
                double mse1 = se(i, j, nl1) + alhc * qa(i, j, nl1);
                mse1 = dmin(mse0, mse1);
                double mss0 = dmax(mse0, mss(i, j, kx));
                int ktop1 = kx, ktop2 = kx;
                for (int k = kx - 3; k >= 3; k--) {
                    double mss2 = mss(i, j, k) + g.wvi[k][2] * (mss(i, j, k + 1) - mss(i, j, k));
                    if (mss0 > mss2) ktop1 = k;
                    if (mse1 > mss2) {
                        ktop2 = k;
                        msthr = mss2;
                    }
                }
                if (ktop1 < kx) {
                    double qthr0 = rhbl * qsat(i, j, kx), qthr1 = rhbl * qsat(i, j, nl1);
                    bool lqthr = (qa(i, j, kx) > qthr0 && qa(i, j, nl1) > qthr1);
                    if (ktop2 < kx) {
                        it = ktop1;
                        qdif(i, j) = dmax(qa(i, j, kx) - qthr0, (mse0 - msthr) * rlhc);
                    } else if (lqthr) {
                        it = ktop1;
                        qdif(i, j) = qa(i, j, kx) - qthr0;
                    }
                }
            }
        }
}

// convection.f90:27-158
void get_convection_tendencies(G2 psa, G3 se, G3 qa, G3 qsat, int *itop, G2 cbmf, G2 precnv, G3 dfse, G3 dfqa,
                               const Geometry &g) {
    const int nl1 = kx - 1, nlp = kx + 1;
    const double fqmax = 5.0;
    Grid2 qdif;
    double entr[kx + 1];
    const double fm0 = p0 * g.dhs[kx] / (grav * trcnv * FL(3600.0));
    const double rdps = 2.0 / (1.0 - psmin);
    for (size_t q = 0; q < NG * kx; q++) dfse.p[q] = 0.0, dfqa.p[q] = 0.0;
    for (size_t q = 0; q < NG; q++) cbmf.p[q] = 0.0, precnv.p[q] = 0.0;
    double sentr = 0.0;
    for (int k = 2; k <= nl1; k++) {
        double e = dmax(0.0, g.fsg[k] - 0.5);
        entr[k] = e * e;
        sentr = sentr + entr[k];
    }
    sentr = entmax / sentr;
    for (int k = 2; k <= nl1; k++) entr[k] = entr[k] * sentr;

    diagnose_convection(psa, se, qa, qsat, itop, qdif, g);

    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            const int it = itop[(i - 1) + ix * (j - 1)];
            if (it == nlp) continue;
            int k = kx, k1 = k - 1;
            double qmax = dmax(FL(1.01) * qa(i, j, k), qsat(i, j, k));
            double sb = se(i, j, k1) + g.wvi[k1][2] * (se(i, j, k) - se(i, j, k1));
            double qb = qa(i, j, k1) + g.wvi[k1][2] * (qa(i, j, k) - qa(i, j, k1));
            qb = dmin(qb, qa(i, j, k));
            double fpsa = psa(i, j) * dmin(1.0, (psa(i, j) - psmin) * rdps);
            double fmass = fm0 * fpsa * dmin(fqmax, qdif(i, j) / (qmax - qb));
            cbmf(i, j) = fmass;
            double fus = fmass * se(i, j, k), fuq = fmass * qmax;
            double fds = fmass * sb, fdq = fmass * qb;
            dfse(i, j, k) = fds - fus;
            dfqa(i, j, k) = fdq - fuq;
            for (k = kx - 1; k >= it + 1; k--) {
                k1 = k - 1;
                dfse(i, j, k) = fus - fds;
                dfqa(i, j, k) = fuq - fdq;
                double enmass = entr[k] * psa(i, j) * cbmf(i, j);
                fmass = fmass + enmass;
                fus = fus + enmass * se(i, j, k);
                fuq = fuq + enmass * qa(i, j, k);
                sb = se(i, j, k1) + g.wvi[k1][2] * (se(i, j, k) - se(i, j, k1));
                qb = qa(i, j, k1) + g.wvi[k1][2] * (qa(i, j, k) - qa(i, j, k1));
                fds = fmass * sb;
                fdq = fmass * qb;
                dfse(i, j, k) = dfse(i, j, k) + fds - fus;
                dfqa(i, j, k) = dfqa(i, j, k) + fdq - fuq;
                double delq = rhil * qsat(i, j, k) - qa(i, j, k);
                if (delq > 0.0) {
                    double fsq = smf * cbmf(i, j) * delq;
                    dfqa(i, j, k) = dfqa(i, j, k) + fsq;
                    dfqa(i, j, kx) = dfqa(i, j, kx) - fsq;
                }
            }
            k = it;
            double qsatb = qsat(i, j, k) + g.wvi[k][2] * (qsat(i, j, k + 1) - qsat(i, j, k));
            precnv(i, j) = dmax(fuq - fmass * qsatb, 0.0);
            dfse(i, j, k) = fus - fds + alhc * precnv(i, j);
            dfqa(i, j, k) = fuq - fdq - precnv(i, j);
        }
}

// ---------------------------------------------------------------------------------------------------
// large_scale_condensation.f90:33-96
void get_large_scale_condensation_tendencies(G2 psa, G3 qa, G3 qsat, int *itop, G2 precls, G3 dtlsc, G3 dqlsc,
                                             const Geometry &g) {
    const double trlsc = FL(4.0), rhlsc = FL(0.9), drhlsc = FL(0.1), rhblsc = FL(0.95);
    const double qsmax = 10.0;
    const double rtlsc = 1.0 / (trlsc * FL(3600.0));
    const double tfact = alhc / cp, prg = p0 / grav;
    Grid2 psa2;
    for (size_t q = 0; q < NG; q++) {
        dtlsc.p[q] = 0.0;
        dqlsc.p[q] = 0.0;
        precls.p[q] = 0.0;
        psa2.d[q] = psa.p[q] * psa.p[q];
    }
    for (int k = 2; k <= kx; k++) {
        double sig2 = g.fsg[k] * g.fsg[k];
        double rhref = rhlsc + drhlsc * (sig2 - 1.0);
        if (k == kx) rhref = dmax(rhref, rhblsc);
        double dqmax = qsmax * sig2 * rtlsc;
        for (int i = 1; i <= ix; i++)
            for (int j = 1; j <= il; j++) {
                double dqa = rhref * qsat(i, j, k) - qa(i, j, k);
                if (dqa < 0.0) {
                    int &it = itop[(i - 1) + ix * (j - 1)];
                    it = std::min(k, it);
                    dqlsc(i, j, k) = dqa * rtlsc;
                    dtlsc(i, j, k) = tfact * dmin(-dqlsc(i, j, k), dqmax * psa2(i, j));
                } else {
                    dqlsc(i, j, k) = 0.0;
                    dtlsc(i, j, k) = 0.0;
                }
            }
    }
    for (int k = 2; k <= kx; k++) {
        double pfact = g.dhs[k] * prg;
        for (size_t q = 0; q < NG; q++) precls.p[q] = precls.p[q] - pfact * dqlsc.p[q + NG * (k - 1)];
    }
    for (size_t q = 0; q < NG; q++) precls.p[q] = precls.p[q] * psa.p[q];
}

// ---------------------------------------------------------------------------------------------------
// shortwave_radiation.f90:14-44
static const double solc = FL(342.0), rhcl1 = FL(0.30), rhcl2 = FL(1.00), qacl = FL(0.20), wpcl = FL(0.2),
                    pmaxcl = FL(10.0), clsmax = FL(0.60), clsminl = FL(0.15), gse_s0 = FL(0.25), gse_s1 = FL(0.40),
                    albcl = FL(0.43), albcls = FL(0.50), epssw = FL(0.020), absdry = FL(0.033), absaer = FL(0.033),
                    abswv1 = FL(0.022), abswv2 = FL(15.000), abscl1 = FL(0.015), abscl2 = FL(0.15),
                    ablwin = FL(0.3), ablwv1 = FL(0.7), ablwv2 = FL(50.0), ablcl1 = FL(12.0), ablcl2 = FL(0.6);

// shortwave_radiation.f90:325-404
void clouds(G3 qa, G3 rh, G2 precnv, G2 precls, const int *iptop, G2 gse, G2 fmask, int *icltop, G2 cloudc,
            G2 clstr, G2 qcloud_equiv) {
    const int nl1 = kx - 1, nlp = kx + 1;
    const double rrcl = 1.0 / (rhcl2 - rhcl1);
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            int &ic = icltop[(i - 1) + ix * (j - 1)];
            if (rh(i, j, nl1) > rhcl1) {
                cloudc(i, j) = rh(i, j, nl1) - rhcl1;
                ic = nl1;
            } else {
                cloudc(i, j) = 0.0;
                ic = nlp;
            }
        }
    for (int k = 3; k <= kx - 2; k++)
        for (int i = 1; i <= ix; i++)
            for (int j = 1; j <= il; j++) {
                double drh = rh(i, j, k) - rhcl1;
                if (drh > cloudc(i, j) && qa(i, j, k) > qacl) {
                    cloudc(i, j) = drh;
                    icltop[(i - 1) + ix * (j - 1)] = k;
                }
            }
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            double pr1 = dmin(pmaxcl, FL(86.4) * (precnv(i, j) + precls(i, j)));
            double c2 = dmin(1.0, cloudc(i, j) * rrcl);
            cloudc(i, j) = dmin(1.0, wpcl * sqrt(pr1) + c2 * c2);
            int &ic = icltop[(i - 1) + ix * (j - 1)];
            ic = std::min(iptop[(i - 1) + ix * (j - 1)], ic);
        }
    for (size_t q = 0; q < NG; q++) qcloud_equiv.p[q] = qa.p[q + NG * (nl1 - 1)];
    const double clfact = FL(1.2);
    const double rgse = 1.0 / (gse_s1 - gse_s0);
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            double fstab = dmax(0.0, dmin(1.0, rgse * (gse(i, j) - gse_s0)));
            clstr(i, j) = fstab * dmax(clsmax - clfact * cloudc(i, j), 0.0);
            double clstrl = dmax(clstr(i, j), clsminl) * rh(i, j, kx);
            clstr(i, j) = clstr(i, j) + fmask(i, j) * (clstrl - clstr(i, j));
        }
}

// shortwave_radiation.f90:50-214
void get_shortwave_rad_fluxes(State &s, G2 psa, G3 qa, const int *icltop, G2 cloudc, G2 clstr) {
    const Geometry &g = s.geo;
    const int nl1 = kx - 1;
    V4<double> tau2{s.p(V_rad_tau2), ix, il, kx};
    V3<double> flux = s.aux3(V_rad_flux), strat = s.aux3(V_rad_strat_corr);
    G3 tt_rsw = s.g3(V_tt_rsw);
    G2 tsr = s.g2(V_tsr), ssrd = s.g2(V_ssrd), ssr = s.g2(V_ssr), zen = s.g2(V_zenit_correction),
       fsol = s.g2(V_flux_solar_in), ozl = s.g2(V_flux_ozone_lower), ozu = s.g2(V_flux_ozone_upper),
       stc = s.g2(V_stratospheric_correction), albs = s.g2(V_alb_surface), qcl = s.g2(V_qcloud_equiv);
    Grid2 acloud, psaz;
    const double fband2 = FL(0.05), fband1 = 1.0 - fband2;
    const double co2 = s.air_absortivity_co2;

    for (size_t q = 0; q < NG * kx * 4; q++) tau2.p[q] = 0.0;
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            int ic = icltop[(i - 1) + ix * (j - 1)];
            if (ic <= kx) tau2(i, j, ic, 3) = albcl * cloudc(i, j);
            tau2(i, j, kx, 3) = albcls * clstr(i, j);
        }
    for (size_t q = 0; q < NG; q++) {
        psaz.d[q] = psa.p[q] * zen.p[q];
        acloud.d[q] = cloudc.p[q] * dmin(abscl1 * qcl.p[q], abscl2);
    }
#define T2(q, k, b) tau2.p[(q) + NG * (((k)-1) + (size_t)kx * ((b)-1))]
#define Q3(a, q, k) a.p[(q) + NG * ((k)-1)]
    for (size_t q = 0; q < NG; q++) T2(q, 1, 1) = exp(-psaz.d[q] * g.dhs[1] * absdry);
    for (int k = 2; k <= nl1; k++) {
        double abs1 = absdry + absaer * (g.fsg[k] * g.fsg[k]);
        for (int i = 1; i <= ix; i++)
            for (int j = 1; j <= il; j++) {
                size_t q = (i - 1) + (size_t)ix * (j - 1);
                if (k >= icltop[q])
                    T2(q, k, 1) = exp(-psaz.d[q] * g.dhs[k] * (abs1 + abswv1 * qa(i, j, k) + acloud.d[q]));
                else
                    T2(q, k, 1) = exp(-psaz.d[q] * g.dhs[k] * (abs1 + abswv1 * qa(i, j, k)));
            }
    }
    {
        double abs1 = absdry + absaer * (g.fsg[kx] * g.fsg[kx]);
        for (size_t q = 0; q < NG; q++) T2(q, kx, 1) = exp(-psaz.d[q] * g.dhs[kx] * (abs1 + abswv1 * Q3(qa, q, kx)));
    }
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) T2(q, k, 2) = exp(-psaz.d[q] * g.dhs[k] * abswv2 * Q3(qa, q, k));

    // 3. downward flux
    for (size_t q = 0; q < NG; q++) {
        tsr.p[q] = fsol.p[q];
        Q3(flux, q, 1) = fsol.p[q] * fband1;
        Q3(flux, q, 2) = fsol.p[q] * fband2;
        // 3.2 stratosphere
        Q3(tt_rsw, q, 1) = Q3(flux, q, 1);
        Q3(flux, q, 1) = T2(q, 1, 1) * (Q3(flux, q, 1) - ozu.p[q] * psa.p[q]);
        Q3(tt_rsw, q, 1) = Q3(tt_rsw, q, 1) - Q3(flux, q, 1);
        Q3(tt_rsw, q, 2) = Q3(flux, q, 1);
        Q3(flux, q, 1) = T2(q, 2, 1) * (Q3(flux, q, 1) - ozl.p[q] * psa.p[q]);
        Q3(tt_rsw, q, 2) = Q3(tt_rsw, q, 2) - Q3(flux, q, 1);
    }
    for (int k = 3; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) {
            T2(q, k, 3) = Q3(flux, q, 1) * T2(q, k, 3);
            Q3(flux, q, 1) = Q3(flux, q, 1) - T2(q, k, 3);
            Q3(tt_rsw, q, k) = Q3(flux, q, 1);
            Q3(flux, q, 1) = T2(q, k, 1) * Q3(flux, q, 1);
            Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) - Q3(flux, q, 1);
        }
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) {
            Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) + Q3(flux, q, 2);
            Q3(flux, q, 2) = T2(q, k, 2) * Q3(flux, q, 2);
            Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) - Q3(flux, q, 2);
        }
    // 4. upward flux
    for (size_t q = 0; q < NG; q++) {
        ssrd.p[q] = Q3(flux, q, 1) + Q3(flux, q, 2);
        Q3(flux, q, 1) = Q3(flux, q, 1) * albs.p[q];
        ssr.p[q] = ssrd.p[q] - Q3(flux, q, 1);
    }
    for (int k = kx; k >= 1; k--)
        for (size_t q = 0; q < NG; q++) {
            Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) + Q3(flux, q, 1);
            Q3(flux, q, 1) = T2(q, k, 1) * Q3(flux, q, 1);
            Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) - Q3(flux, q, 1);
            Q3(flux, q, 1) = Q3(flux, q, 1) + T2(q, k, 3);
        }
    for (size_t q = 0; q < NG; q++) tsr.p[q] = tsr.p[q] - Q3(flux, q, 1);

    // 5. longwave transmissivities
    for (size_t q = 0; q < NG; q++) {
        T2(q, 1, 1) = exp(-psa.p[q] * g.dhs[1] * ablwin);
        T2(q, 1, 2) = exp(-psa.p[q] * g.dhs[1] * co2);
        T2(q, 1, 3) = 1.0;
        T2(q, 1, 4) = 1.0;
    }
    for (int k = 2; k <= kx; k += kx - 2)
        for (size_t q = 0; q < NG; q++) {
            T2(q, k, 1) = exp(-psa.p[q] * g.dhs[k] * ablwin);
            T2(q, k, 2) = exp(-psa.p[q] * g.dhs[k] * co2);
            T2(q, k, 3) = exp(-psa.p[q] * g.dhs[k] * ablwv1 * Q3(qa, q, k));
            T2(q, k, 4) = exp(-psa.p[q] * g.dhs[k] * ablwv2 * Q3(qa, q, k));
        }
    for (size_t q = 0; q < NG; q++) acloud.d[q] = cloudc.p[q] * ablcl2;
    for (int k = 3; k <= nl1; k++)
        for (size_t q = 0; q < NG; q++) {
            double deltap = psa.p[q] * g.dhs[k];
            double acloud1 = (k < icltop[q]) ? acloud.d[q] : ablcl1 * cloudc.p[q];
            T2(q, k, 1) = exp(-deltap * (ablwin + acloud1));
            T2(q, k, 2) = exp(-deltap * co2);
            T2(q, k, 3) = exp(-deltap * dmax(ablwv1 * Q3(qa, q, k), acloud.d[q]));
            T2(q, k, 4) = exp(-deltap * dmax(ablwv2 * Q3(qa, q, k), acloud.d[q]));
        }
    const double eps1 = epslw / (g.dhs[1] + g.dhs[2]);
    for (size_t q = 0; q < NG; q++) {
        Q3(strat, q, 1) = stc.p[q] * psa.p[q];
        Q3(strat, q, 2) = eps1 * psa.p[q];
    }
}

// shortwave_radiation.f90:276-322
static void solar(double tyear, double csol, double *topsr /*1-based il*/, const Geometry &g) {
    const double pigr = 2.0 * F_ASIN1;  // REAL(4) 2.0*asin(1.0): exact doubling
    const double alpha = 2.0 * pigr * tyear;
    const double ca1 = cos(alpha), sa1 = sin(alpha);
    const double ca2 = ca1 * ca1 - sa1 * sa1, sa2 = 2.0 * sa1 * ca1;
    const double ca3 = ca1 * ca2 - sa1 * sa2, sa3 = sa1 * ca2 + sa2 * ca1;
    const double decl = FL(0.006918) - FL(0.399912) * ca1 + FL(0.070257) * sa1 - FL(0.006758) * ca2 +
                        FL(0.000907) * sa2 - FL(0.002697) * ca3 + FL(0.001480) * sa3;
    const double fdis = FL(1.000110) + FL(0.034221) * ca1 + FL(0.001280) * sa1 + FL(0.000719) * ca2 + FL(0.000077) * sa2;
    const double cdecl = cos(decl), sdecl = sin(decl), tdecl = sdecl / cdecl;
    const double csolp = csol / pigr;
    for (int j = 1; j <= il; j++) {
        double ch0 = dmin(1.0, dmax(-1.0, -tdecl * g.sia[j] / g.coa[j]));
        double h0 = acos(ch0), sh0 = sin(h0);
        topsr[j] = csolp * fdis * (h0 * g.sia[j] * sdecl + sh0 * g.coa[j] * cdecl);
    }
}

// shortwave_radiation.f90:218-273
void get_zonal_average_fields(State &s, double tyear) {
    const Geometry &g = s.geo;
    double topsr[il + 1];
    // 4.0*asin(1.0) and 10.0/365.0 are REAL(4) constant expressions
    const double alpha = (double)(4.0f * (float)F_ASIN1) * (tyear + (double)(10.0f / 365.0f));
    const double dalpha = 0.0;
    const double coz1 = 1.0 * dmax(0.0, cos(alpha - dalpha));
    const double coz2 = FL(1.8), azen = 1.0, nzen = 2.0;
    const double rzen = -cos(alpha) * FL(23.45) * F_ASIN1 / FL(90.0);
    const double fs0 = 6.0;
    solar(tyear, 4.0 * solc, topsr, g);
    G2 fsol = s.g2(V_flux_solar_in), ozl = s.g2(V_flux_ozone_lower), ozu = s.g2(V_flux_ozone_upper),
       zen = s.g2(V_zenit_correction), stc = s.g2(V_stratospheric_correction);
    for (int j = 1; j <= il; j++) {
        double flat2 = FL(1.5) * (g.sia[j] * g.sia[j]) - 0.5;
        double ou = 0.5 * epssw;
        double ol = FL(0.4) * epssw * (1.0 + coz1 * g.sia[j] + coz2 * flat2);
        double zc = 1.0 + azen * pow(1.0 - (g.coa[j] * cos(rzen) + g.sia[j] * sin(rzen)), nzen);
        for (int i = 1; i <= ix; i++) {
            fsol(i, j) = topsr[j];
            zen(i, j) = zc;
            ozu(i, j) = fsol(i, j) * ou * zc;
            ozl(i, j) = fsol(i, j) * ol * zc;
            stc(i, j) = dmax(fs0 - fsol(i, j), 0.0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// longwave_radiation.f90:208-232 ; fband(100:400, 4)
#define FB(t, jb) fband[((t)-100) + 301 * ((jb)-1)]
void radset(double *fband) {
    const double eps1 = 1.0 - epslw;
    for (int jt = 200; jt <= 320; jt++) {
        // polynomials evaluated in REAL(4): literal - literal*float(int**2)
        FB(jt, 2) = (double)(0.148f - 3.0e-6f * (float)((jt - 247) * (jt - 247))) * eps1;
        FB(jt, 3) = (double)(0.356f - 5.2e-6f * (float)((jt - 282) * (jt - 282))) * eps1;
        FB(jt, 4) = (double)(0.314f + 1.0e-5f * (float)((jt - 315) * (jt - 315))) * eps1;
        FB(jt, 1) = eps1 - (FB(jt, 2) + FB(jt, 3) + FB(jt, 4));
    }
    for (int jb = 1; jb <= 4; jb++) {
        for (int jt = 100; jt <= 199; jt++) FB(jt, jb) = FB(200, jb);
        for (int jt = 321; jt <= 400; jt++) FB(jt, jb) = FB(320, jb);
    }
}
static inline double fband_at(const double *fband, double t, int jb) {
    long it = nint(t);
    // the reference indexes fband(100:400) unguarded (UB outside); clamp, documented in DESIGN.md
    if (it < 100) it = 100;
    if (it > 400) it = 400;
    return FB(it, jb);
}

// longwave_radiation.f90:16-121
void get_downward_longwave_rad_fluxes(State &s, G3 ta, G2 fsfcd, G3 dfabs) {
    const Geometry &g = s.geo;
    const int nl1 = kx - 1, nband = 4;
    const double *fband = s.p(V_fband);
    V4<double> tau2{s.p(V_rad_tau2), ix, il, kx};
    V3<double> flux = s.aux3(V_rad_flux);
    V4<double> st4a{s.p(V_rad_st4a), ix, il, kx};
#define S4(q, k, b) st4a.p[(q) + NG * (((k)-1) + (size_t)kx * ((b)-1))]
    for (int k = 1; k <= nl1; k++)
        for (size_t q = 0; q < NG; q++) S4(q, k, 1) = Q3(ta, q, k) + g.wvi[k][2] * (Q3(ta, q, k + 1) - Q3(ta, q, k));
    for (size_t q = 0; q < NG; q++) {
        S4(q, 1, 2) = FL(0.75) * Q3(ta, q, 1) + FL(0.25) * S4(q, 1, 1);
        S4(q, 2, 2) = FL(0.50) * Q3(ta, q, 2) + FL(0.25) * (S4(q, 1, 1) + S4(q, 2, 1));
    }
    const double anis = 1.0;
    for (int k = 3; k <= nl1; k++)
        for (size_t q = 0; q < NG; q++) S4(q, k, 2) = 0.5 * anis * dmax(S4(q, k, 1) - S4(q, k - 1, 1), 0.0);
    for (size_t q = 0; q < NG; q++) S4(q, kx, 2) = anis * dmax(Q3(ta, q, kx) - S4(q, nl1, 1), 0.0);
    for (int k = 1; k <= 2; k++)
        for (size_t q = 0; q < NG; q++) {
            S4(q, k, 1) = sbc * pow(S4(q, k, 2), 4.0);
            S4(q, k, 2) = 0.0;
        }
    for (int k = 3; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) {
            double st3a = sbc * pow(Q3(ta, q, k), 3.0);
            S4(q, k, 1) = st3a * Q3(ta, q, k);
            S4(q, k, 2) = 4.0 * st3a * S4(q, k, 2);
        }
    for (size_t q = 0; q < NG; q++) fsfcd.p[q] = 0.0;
    for (size_t q = 0; q < NG * kx; q++) dfabs.p[q] = 0.0;
    {
        const int k = 1;
        for (int jb = 1; jb <= 2; jb++)
            for (size_t q = 0; q < NG; q++) {
                double emis = 1.0 - T2(q, k, jb);
                double brad = fband_at(fband, Q3(ta, q, k), jb) * (S4(q, k, 1) + emis * S4(q, k, 2));
                Q3(flux, q, jb) = emis * brad;
                Q3(dfabs, q, k) = Q3(dfabs, q, k) - Q3(flux, q, jb);
            }
    }
    for (int jb = 3; jb <= nband; jb++)
        for (size_t q = 0; q < NG; q++) Q3(flux, q, jb) = 0.0;
    for (int jb = 1; jb <= nband; jb++)
        for (int k = 2; k <= kx; k++)
            for (size_t q = 0; q < NG; q++) {
                double emis = 1.0 - T2(q, k, jb);
                double brad = fband_at(fband, Q3(ta, q, k), jb) * (S4(q, k, 1) + emis * S4(q, k, 2));
                Q3(dfabs, q, k) = Q3(dfabs, q, k) + Q3(flux, q, jb);
                Q3(flux, q, jb) = T2(q, k, jb) * Q3(flux, q, jb) + emis * brad;
                Q3(dfabs, q, k) = Q3(dfabs, q, k) - Q3(flux, q, jb);
            }
    for (int jb = 1; jb <= nband; jb++)
        for (size_t q = 0; q < NG; q++) fsfcd.p[q] = fsfcd.p[q] + emisfc * Q3(flux, q, jb);
    for (size_t q = 0; q < NG; q++) {
        double corlw = epslw * emisfc * S4(q, kx, 1);
        Q3(dfabs, q, kx) = Q3(dfabs, q, kx) - corlw;
        fsfcd.p[q] = fsfcd.p[q] + corlw;
    }
}

// longwave_radiation.f90:124-205
void get_upward_longwave_rad_fluxes(State &s, G3 ta, G2 ts, G2 fsfcd, G2 fsfcu, G2 fsfc, G2 ftop, G3 dfabs) {
    const Geometry &g = s.geo;
    const int nband = 4;
    const double *fband = s.p(V_fband);
    V4<double> tau2{s.p(V_rad_tau2), ix, il, kx};
    V3<double> flux = s.aux3(V_rad_flux), strat = s.aux3(V_rad_strat_corr);
    V4<double> st4a{s.p(V_rad_st4a), ix, il, kx};
    const double refsfc = 1.0 - emisfc;
    for (size_t q = 0; q < NG; q++) fsfc.p[q] = fsfcu.p[q] - fsfcd.p[q];
    for (int jb = 1; jb <= nband; jb++)
        for (size_t q = 0; q < NG; q++)
            Q3(flux, q, jb) = fband_at(fband, ts.p[q], jb) * fsfcu.p[q] + refsfc * Q3(flux, q, jb);
    for (size_t q = 0; q < NG; q++) Q3(dfabs, q, kx) = Q3(dfabs, q, kx) + epslw * fsfcu.p[q];
    for (int jb = 1; jb <= nband; jb++)
        for (int k = kx; k >= 2; k--)
            for (size_t q = 0; q < NG; q++) {
                double emis = 1.0 - T2(q, k, jb);
                double brad = fband_at(fband, Q3(ta, q, k), jb) * (S4(q, k, 1) - emis * S4(q, k, 2));
                Q3(dfabs, q, k) = Q3(dfabs, q, k) + Q3(flux, q, jb);
                Q3(flux, q, jb) = T2(q, k, jb) * Q3(flux, q, jb) + emis * brad;
                Q3(dfabs, q, k) = Q3(dfabs, q, k) - Q3(flux, q, jb);
            }
    {
        const int k = 1;
        for (int jb = 1; jb <= 2; jb++)
            for (size_t q = 0; q < NG; q++) {
                double emis = 1.0 - T2(q, k, jb);
                double brad = fband_at(fband, Q3(ta, q, k), jb) * (S4(q, k, 1) - emis * S4(q, k, 2));
                Q3(dfabs, q, k) = Q3(dfabs, q, k) + Q3(flux, q, jb);
                Q3(flux, q, jb) = T2(q, k, jb) * Q3(flux, q, jb) + emis * brad;
                Q3(dfabs, q, k) = Q3(dfabs, q, k) - Q3(flux, q, jb);
            }
    }
    for (size_t q = 0; q < NG; q++) {
        double corlw1 = g.dhs[1] * Q3(strat, q, 2) * S4(q, 1, 1) + Q3(strat, q, 1);
        double corlw2 = g.dhs[2] * Q3(strat, q, 2) * S4(q, 2, 1);
        Q3(dfabs, q, 1) = Q3(dfabs, q, 1) - corlw1;
        Q3(dfabs, q, 2) = Q3(dfabs, q, 2) - corlw2;
        ftop.p[q] = corlw1 + corlw2;
    }
    for (int jb = 1; jb <= nband; jb++)
        for (size_t q = 0; q < NG; q++) ftop.p[q] = ftop.p[q] + Q3(flux, q, jb);
}

// ---------------------------------------------------------------------------------------------------
// surface_fluxes.f90:13-32
static const double fwind0 = FL(0.95), ftemp0 = FL(1.0), fhum0 = FL(0.0), cdl = FL(2.4e-3), cds = FL(1.0e-3),
                    chl = FL(1.2e-3), chs = FL(0.9e-3), vgust = FL(5.0), ctday = FL(1.0e-2), dtheta = FL(3.0),
                    fstab_sf = FL(0.67), hdrag = FL(2000.0), clambda = FL(7.0), clambsn = FL(7.0);

// surface_fluxes.f90:40-320  (lfluxland = .true.; the second call in physics.f90:188 is dead: sea_coupling_flag=0)
void get_surface_fluxes(State &s, G2 psa, G3 ua, G3 va, G3 ta, G3 qa, G3 rh, G3 phi, G2 tsea, G2 tsfc, G2 tskin,
                        G2 u0, G2 v0, G2 t0) {
    (void)rh;
    const Geometry &g = s.geo;
    G2 phi0 = s.g2(V_phis0), fmask = s.g2(V_fmask_land), forog = s.g2(V_forog), ssrd = s.g2(V_ssrd),
       slrd = s.g2(V_slrd), alb_land = s.g2(V_alb_land), alb_sea = s.g2(V_alb_sea), snowc = s.g2(V_snowc),
       land_temp = s.g2(V_land_temp), saw = s.g2(V_soil_avail_water);
    V3<double> ustr = s.aux3(V_ustr), vstr = s.aux3(V_vstr), shf = s.aux3(V_shf), evap = s.aux3(V_evap),
               slru = s.aux3(V_slru), hfluxn = s.aux3(V_hfluxn);
    Grid3 t1(2), q1(2), t2(2), qsat0(2), denvvs(3);  // denvvs(:,:,0:2) -> slabs 1..3
    Grid2 dslr, dtskin, clamb, cdsdv, tsk3;
    const int nl1 = kx - 1;
    const double esbc = emisfc * sbc;
    const double ghum0 = 1.0 - fhum0;
    (void)ghum0;
    for (size_t q = 0; q < NG; q++) {
        u0.p[q] = fwind0 * Q3(ua, q, kx);
        v0.p[q] = fwind0 * Q3(va, q, kx);
    }
    const double gtemp0 = 1.0 - ftemp0, rcp = 1.0 / cp;
    for (size_t q = 0; q < NG; q++) {
        double dt1 = g.wvi[kx][2] * (Q3(ta, q, kx) - Q3(ta, q, nl1));
        Q3(t1, q, 1) = Q3(ta, q, kx) + dt1;
        Q3(t1, q, 2) = Q3(t1, q, 1) - phi0.p[q] * dt1 / (rgas * FL(288.0) * g.sigl[kx]);
        Q3(t2, q, 2) = Q3(ta, q, kx) + rcp * Q3(phi, q, kx);
        Q3(t2, q, 1) = Q3(t2, q, 2) - rcp * phi0.p[q];
    }
    for (size_t q = 0; q < NG; q++) {
        if (Q3(ta, q, kx) > Q3(ta, q, nl1)) {
            Q3(t1, q, 1) = ftemp0 * Q3(t1, q, 1) + gtemp0 * Q3(t2, q, 1);
            Q3(t1, q, 2) = ftemp0 * Q3(t1, q, 2) + gtemp0 * Q3(t2, q, 2);
        } else {
            Q3(t1, q, 1) = Q3(ta, q, kx);
            Q3(t1, q, 2) = Q3(ta, q, kx);
        }
        t0.p[q] = Q3(t1, q, 2) + fmask.p[q] * (Q3(t1, q, 1) - Q3(t1, q, 2));
    }
    for (size_t q = 0; q < NG; q++)
        Q3(denvvs, q, 1) = (p0 * psa.p[q] / (rgas * t0.p[q])) * sqrt(u0.p[q] * u0.p[q] + v0.p[q] * v0.p[q] + vgust * vgust);
    for (int j = 1; j <= il; j++)
        for (int i = 1; i <= ix; i++)
            tskin(i, j) = land_temp(i, j) + ctday * sqrt(g.coa[j]) * ssrd(i, j) * (1.0 - alb_land(i, j)) * psa(i, j);
    double rdth = fstab_sf / dtheta;
    double astab = 0.5;  // lscasym = .true.
    for (size_t q = 0; q < NG; q++) {
        double dthl;
        if (tskin.p[q] > Q3(t2, q, 1))
            dthl = dmin(dtheta, tskin.p[q] - Q3(t2, q, 1));
        else
            dthl = dmax(-dtheta, astab * (tskin.p[q] - Q3(t2, q, 1)));
        Q3(denvvs, q, 2) = Q3(denvvs, q, 1) * (1.0 + dthl * rdth);
    }
    for (size_t q = 0; q < NG; q++) {
        double cdldv = cdl * Q3(denvvs, q, 1) * forog.p[q];
        Q3(ustr, q, 1) = -cdldv * Q3(ua, q, kx);
        Q3(vstr, q, 1) = -cdldv * Q3(va, q, kx);
    }
    const double chlcp = chl * cp;
    for (size_t q = 0; q < NG; q++) Q3(shf, q, 1) = chlcp * Q3(denvvs, q, 2) * (tskin.p[q] - Q3(t1, q, 1));
    for (size_t q = 0; q < NG; q++) Q3(q1, q, 1) = Q3(qa, q, kx);  // fhum0 = 0 branch
    get_qsat(tskin.p, psa.p, 1.0, qsat0.d.data(), (int)NG);
    for (size_t q = 0; q < NG; q++)
        Q3(evap, q, 1) = chl * Q3(denvvs, q, 2) * dmax(0.0, saw.p[q] * Q3(qsat0, q, 1) - Q3(q1, q, 1));
    for (size_t q = 0; q < NG; q++) {
        tsk3.d[q] = pow(tskin.p[q], 3.0);
        dslr.d[q] = 4.0 * esbc * tsk3.d[q];
        Q3(slru, q, 1) = esbc * tsk3.d[q] * tskin.p[q];
        Q3(hfluxn, q, 1) = ssrd.p[q] * (1.0 - alb_land.p[q]) + slrd.p[q] -
                           (Q3(slru, q, 1) + Q3(shf, q, 1) + alhc * Q3(evap, q, 1));
    }
    // lskineb = .true.
    for (size_t q = 0; q < NG; q++) {
        clamb.d[q] = clambda + snowc.p[q] * (clambsn - clambda);
        Q3(hfluxn, q, 1) = Q3(hfluxn, q, 1) - clamb.d[q] * (tskin.p[q] - land_temp.p[q]);
        dtskin.d[q] = tskin.p[q] + 1.0;
    }
    get_qsat(dtskin.d.data(), psa.p, 1.0, qsat0.d.data() + NG, (int)NG);
    for (size_t q = 0; q < NG; q++) {
        if (Q3(evap, q, 1) > 0.0)
            Q3(qsat0, q, 2) = saw.p[q] * (Q3(qsat0, q, 2) - Q3(qsat0, q, 1));
        else
            Q3(qsat0, q, 2) = 0.0;
    }
    for (size_t q = 0; q < NG; q++) {
        dtskin.d[q] = Q3(hfluxn, q, 1) / (clamb.d[q] + dslr.d[q] + chl * Q3(denvvs, q, 2) * (cp + alhc * Q3(qsat0, q, 2)));
        tskin.p[q] = tskin.p[q] + dtskin.d[q];
        Q3(shf, q, 1) = Q3(shf, q, 1) + chlcp * Q3(denvvs, q, 2) * dtskin.d[q];
        Q3(evap, q, 1) = Q3(evap, q, 1) + chl * Q3(denvvs, q, 2) * Q3(qsat0, q, 2) * dtskin.d[q];
        Q3(slru, q, 1) = Q3(slru, q, 1) + dslr.d[q] * dtskin.d[q];
        Q3(hfluxn, q, 1) = clamb.d[q] * (tskin.p[q] - land_temp.p[q]);
    }
    rdth = fstab_sf / dtheta;
    for (size_t q = 0; q < NG; q++) {
        double dths;
        if (tsea.p[q] > Q3(t2, q, 2))
            dths = dmin(dtheta, tsea.p[q] - Q3(t2, q, 2));
        else
            dths = dmax(-dtheta, astab * (tsea.p[q] - Q3(t2, q, 2)));
        Q3(denvvs, q, 3) = Q3(denvvs, q, 1) * (1.0 + dths * rdth);
    }
    for (size_t q = 0; q < NG; q++) Q3(q1, q, 2) = Q3(qa, q, kx);
    for (size_t q = 0; q < NG; q++) {
        cdsdv.d[q] = cds * Q3(denvvs, q, 3);
        Q3(ustr, q, 2) = -cdsdv.d[q] * Q3(ua, q, kx);
        Q3(vstr, q, 2) = -cdsdv.d[q] * Q3(va, q, kx);
    }
    // sea surface
    for (size_t q = 0; q < NG; q++) Q3(shf, q, 2) = chs * cp * Q3(denvvs, q, 3) * (tsea.p[q] - Q3(t1, q, 2));
    get_qsat(tsea.p, psa.p, 1.0, qsat0.d.data() + NG, (int)NG);
    for (size_t q = 0; q < NG; q++) {
        Q3(evap, q, 2) = chs * Q3(denvvs, q, 3) * (Q3(qsat0, q, 2) - Q3(q1, q, 2));
        Q3(slru, q, 2) = esbc * pow(tsea.p[q], 4.0);
        Q3(hfluxn, q, 2) = ssrd.p[q] * (1.0 - alb_sea.p[q]) + slrd.p[q] - Q3(slru, q, 2) + Q3(shf, q, 2) +
                           alhc * Q3(evap, q, 2);
    }
    for (size_t q = 0; q < NG; q++) {
        Q3(ustr, q, 3) = Q3(ustr, q, 2) + fmask.p[q] * (Q3(ustr, q, 1) - Q3(ustr, q, 2));
        Q3(vstr, q, 3) = Q3(vstr, q, 2) + fmask.p[q] * (Q3(vstr, q, 1) - Q3(vstr, q, 2));
        Q3(shf, q, 3) = Q3(shf, q, 2) + fmask.p[q] * (Q3(shf, q, 1) - Q3(shf, q, 2));
        Q3(evap, q, 3) = Q3(evap, q, 2) + fmask.p[q] * (Q3(evap, q, 1) - Q3(evap, q, 2));
        Q3(slru, q, 3) = Q3(slru, q, 2) + fmask.p[q] * (Q3(slru, q, 1) - Q3(slru, q, 2));
        tsfc.p[q] = tsea.p[q] + fmask.p[q] * (land_temp.p[q] - tsea.p[q]);
        tskin.p[q] = tsea.p[q] + fmask.p[q] * (tskin.p[q] - tsea.p[q]);
        t0.p[q] = Q3(t1, q, 2) + fmask.p[q] * (Q3(t1, q, 1) - Q3(t1, q, 2));
    }
}

// surface_fluxes.f90:324-334
void set_orog_land_sfc_drag(G2 phi0, G2 forog) {
    const double rhdrag = 1.0 / (grav * hdrag);
    for (size_t q = 0; q < NG; q++) forog.p[q] = 1.0 + rhdrag * (1.0 - exp(-dmax(phi0.p[q], 0.0) * rhdrag));
}

// ---------------------------------------------------------------------------------------------------
// vertical_diffusion.f90:30-146
void get_vertical_diffusion_tend(G3 se, G3 rh, G3 qa, G3 qsat, G3 phi, const int *icnv, G3 utenvd, G3 vtenvd,
                                 G3 ttenvd, G3 qtenvd, const Geometry &g) {
    const double trshc = FL(6.0), trvdi = FL(24.0), trvds = FL(6.0), redshc = FL(0.5), rhgrad = FL(0.5), segrad = FL(0.1);
    const int nl1 = kx - 1;
    double rsig[kx + 1], rsig1[kx + 1];
    const double cshc = g.dhs[kx] / FL(3600.0);
    const double cvdi = (g.sigh[nl1] - g.sigh[1]) / (double)((float)(nl1 - 1) * 3600.0f);
    const double fshcq = cshc / trshc, fshcse = cshc / (trshc * cp);
    const double fvdiq = cvdi / trvdi, fvdise = cvdi / (trvds * cp);
    for (int k = 1; k <= nl1; k++) {
        rsig[k] = 1.0 / g.dhs[k];
        rsig1[k] = 1.0 / (1.0 - g.sigh[k]);
    }
    rsig[kx] = 1.0 / g.dhs[kx];
    for (size_t q = 0; q < NG * kx; q++) utenvd.p[q] = 0.0, vtenvd.p[q] = 0.0, ttenvd.p[q] = 0.0, qtenvd.p[q] = 0.0;
    double drh0 = rhgrad * (g.fsg[kx] - g.fsg[nl1]);
    double fvdiq2 = fvdiq * g.sigh[nl1];
    for (size_t q = 0; q < NG; q++) {
        double dmse = Q3(se, q, kx) - Q3(se, q, nl1) + alhc * (Q3(qa, q, kx) - Q3(qsat, q, nl1));
        double drh = Q3(rh, q, kx) - Q3(rh, q, nl1);
        double fcnv = 1.0;
        if (dmse >= 0.0) {
            if (icnv[q] > 0) fcnv = redshc;
            double fluxse = fcnv * fshcse * dmse;
            Q3(ttenvd, q, nl1) = fluxse * rsig[nl1];
            Q3(ttenvd, q, kx) = -fluxse * rsig[kx];
            if (drh >= 0.0) {
                double fluxq = fcnv * fshcq * Q3(qsat, q, kx) * drh;
                Q3(qtenvd, q, nl1) = fluxq * rsig[nl1];
                Q3(qtenvd, q, kx) = -fluxq * rsig[kx];
            }
        } else if (drh > drh0) {
            double fluxq = fvdiq2 * Q3(qsat, q, nl1) * drh;
            Q3(qtenvd, q, nl1) = fluxq * rsig[nl1];
            Q3(qtenvd, q, kx) = -fluxq * rsig[kx];
        }
    }
    for (int k = 3; k <= kx - 2; k++)
        if (g.sigh[k] > 0.5) {
            drh0 = rhgrad * (g.fsg[k + 1] - g.fsg[k]);
            fvdiq2 = fvdiq * g.sigh[k];
            for (size_t q = 0; q < NG; q++) {
                double drh = Q3(rh, q, k + 1) - Q3(rh, q, k);
                if (drh >= drh0) {
                    double fluxq = fvdiq2 * Q3(qsat, q, k) * drh;
                    Q3(qtenvd, q, k) = Q3(qtenvd, q, k) + fluxq * rsig[k];
                    Q3(qtenvd, q, k + 1) = Q3(qtenvd, q, k + 1) - fluxq * rsig[k + 1];
                }
            }
        }
    for (int k = 1; k <= nl1; k++)
        for (size_t q = 0; q < NG; q++) {
            double se0 = Q3(se, q, k + 1) + segrad * (Q3(phi, q, k) - Q3(phi, q, k + 1));
            if (Q3(se, q, k) < se0) {
                double fluxse = fvdise * (se0 - Q3(se, q, k));
                Q3(ttenvd, q, k) = Q3(ttenvd, q, k) + fluxse * rsig[k];
                for (int k1 = k + 1; k1 <= kx; k1++) Q3(ttenvd, q, k1) = Q3(ttenvd, q, k1) - fluxse * rsig1[k];
            }
        }
}

// ---------------------------------------------------------------------------------------------------
// physics.f90:14-101 : spectral -> grid part, then the column physics proper
// ---------------------------------------------------------------------------------------------------
// sppt.f90:40-146 -- stochastically perturbed parametrisation tendencies (Palmer et al. 2009).
// The reference routine cannot run as written: `sigma` is allocated (ix,il,kx) and assigned an (mx,nx) array, the AR(1)
// state `sppt_spec` is a local allocatable that is freed at every return (so `phi * sppt_spec` reads unallocated memory),
// the result is deallocated before it is returned, and the generator is seeded from the system clock.  This restates the
// ALGORITHM those lines describe (the one of the routine's origin, speedy.f90 by S. Hatfield, where sigma and sppt_spec
// are (mx,nx,kx) module arrays): complex Gaussian noise clipped to +-10, sigma(m,n) = f0 exp(-L^2 el2 / 4), AR(1) with
// phi = exp(-(24/nsteps)/6 h), spec2grid, clipping to +-1 -- with the pattern kept per member and a counter-based
// generator (splitmix64) in place of random_number, so that runs are reproducible and the GPU kernel can be compared.
uint64_t sppt_mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// sppt.f90:118-133 randn(0, 1): Box-Muller with the reference's constants ((-2 log r1)**0.5, v = 2.0*6.28318530718 * r2
// in REAL(4), sin); r1 in (0, 1], r2 in [0, 1) from the counter `ctr`
static double sppt_randn(uint64_t key, uint64_t ctr) {
    const double r1 = ((double)(sppt_mix64(key + 2 * ctr) >> 11) + 1.0) * 0x1p-53;
    const double r2 = (double)(sppt_mix64(key + 2 * ctr + 1) >> 11) * 0x1p-53;
    const double u = pow(-2.0 * log(r1), 0.5);
    const double v = (double)(2.0f * 6.28318530718f) * r2;
    return 0.0 + 1.0 * u * sin(v);
}
void gen_sppt(State &s, G3 sppt_grid) {
    const double time_decorr = 6.0, len_decorr = 500000.0, stddev = FL(0.33);  // :27-37
    const double phi = exp(-(24 / (double)nsteps) / time_decorr);
    if (s.sppt_spec.empty()) s.sppt_spec.assign((size_t)mx * nx * kx, cplx{0.0, 0.0});
    const uint64_t key = sppt_mix64(s.sppt_seed ^ sppt_mix64(s.sppt_member));
    double f0 = 0.0;  // :87-88
    for (int n = 1; n <= trunc_; n++) f0 = f0 + (2 * n + 1) * exp(-0.5 * ((len_decorr / rearth) * (len_decorr / rearth)) * n * (n + 1));
    f0 = sqrt((stddev * stddev * (1 - phi * phi)) / (2 * f0));
    const double first_fac = pow(1 - phi * phi, -0.5);
    for (int k = 1; k <= kx; k++)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) {
                const uint64_t ctr = (s.sppt_calls * (uint64_t)(mx * nx) + (uint64_t)((m - 1) + mx * (n - 1))) * kx + (k - 1);
                const double rr = sppt_randn(key, 2 * ctr), ri = sppt_randn(key, 2 * ctr + 1);
                const cplx eta = {fmin(10.0, fabs(rr)) * copysign(1.0, rr), fmin(10.0, fabs(ri)) * copysign(1.0, ri)};  // :73-76
                const double sigma = f0 * exp(-0.25 * (len_decorr * len_decorr) * s.spec.el2[Spectral::i2(m, n)]);     // :90-92
                cplx &x = s.sppt_spec[(m - 1) + (size_t)mx * ((n - 1) + (size_t)nx * (k - 1))];
                if (s.sppt_calls == 0) x = (first_fac * sigma) * eta;  // :95
                else x = phi * x + sigma * eta;                   // :100
            }
    s.sppt_calls += 1;
    for (int k = 1; k <= kx; k++)
        s.spec.spec2grid(S2{s.sppt_spec.data() + (size_t)mx * nx * (k - 1), mx}, sppt_grid.slab(k), 1);
    for (size_t q = 0; q < NG * kx; q++) sppt_grid.p[q] = fmin(1.0, fabs(sppt_grid.p[q])) * copysign(1.0, sppt_grid.p[q]);  // :109
}

void get_physical_tendencies(State &s, int j1, G3 utend, G3 vtend, G3 ttend, G3 qtend) {
    const Spectral &sp = s.spec;
    Grid3 ug, vg, tg, qg, phig;
    Grid3 utend_dyn, vtend_dyn, ttend_dyn, qtend_dyn;  // physics.f90:80-83
    if (s.sppt_on)
        for (size_t q = 0; q < NG * kx; q++)
            utend_dyn.d[q] = utend.p[q], vtend_dyn.d[q] = vtend.p[q], ttend_dyn.d[q] = ttend.p[q], qtend_dyn.d[q] = qtend.p[q];
    Spec2 ucos, vcos;
    Grid2 pslg;
    S3 vor = s.s4lev(V_vor, j1), div = s.s4lev(V_div, j1), t = s.s4lev(V_t, j1), tr = s.s4lev(V_tr, j1),
       phi = s.s3(V_phi);
    for (int k = 1; k <= kx; k++) {  // :89-99
        sp.vort2vel(vor.slab(k), div.slab(k), ucos, vcos);
        sp.spec2grid(ucos, ug.slab(k), 2);
        sp.spec2grid(vcos, vg.slab(k), 2);
        sp.spec2grid(t.slab(k), tg.slab(k), 1);
        sp.spec2grid(tr.slab(k), qg.slab(k), 1);
        sp.spec2grid(phi.slab(k), phig.slab(k), 1);
    }
    sp.spec2grid(s.s3lev(V_ps, j1), pslg, 1);
    physics_columns(s, ug, vg, tg, qg, phig, pslg, utend, vtend, ttend, qtend, nullptr);
    if (s.sppt_on) {  // physics.f90:233-248 (mu(k) = 1, sppt.f90:20)
        Grid3 pat;
        gen_sppt(s, pat);
        s.sppt_last = pat.d;
        const double mu = 1.0;
        for (size_t q = 0; q < NG * kx; q++) {
            utend.p[q] = (1 + pat.d[q] * mu) * (utend.p[q] - utend_dyn.d[q]) + utend_dyn.d[q];
            vtend.p[q] = (1 + pat.d[q] * mu) * (vtend.p[q] - vtend_dyn.d[q]) + vtend_dyn.d[q];
            ttend.p[q] = (1 + pat.d[q] * mu) * (ttend.p[q] - ttend_dyn.d[q]) + ttend_dyn.d[q];
            qtend.p[q] = (1 + pat.d[q] * mu) * (qtend.p[q] - qtend_dyn.d[q]) + qtend_dyn.d[q];
        }
    }
}

// physics.f90:103-231 : grid-point (column-independent) part.  qg is modified in place (max(qg,0)).
// `dbg` (optional, 3*NG ints) receives iptop (after LSC), icnv and icltop for exact index checks.
void physics_columns(State &s, G3 ug, G3 vg, G3 tg, G3 qg, G3 phig, G2 pslg, G3 utend, G3 vtend, G3 ttend,
                     G3 qtend, int *dbg) {
    const Geometry &g = s.geo;
    Grid3 tt_cnv, qt_cnv, tt_lsc, qt_lsc, tt_rlw, ut_pbl, vt_pbl, tt_pbl, qt_pbl;
    Grid3 se, rh, qsat;
    Grid2 rps, gse, psg, ts, tskin, u0, v0, t0, cloudc, clstr;
    std::vector<int> iptop(NG), icnv(NG), icltop(NG, 0);
    for (size_t q = 0; q < NG; q++) {
        psg.d[q] = exp(pslg.p[q]);
        rps.d[q] = 1.0 / psg.d[q];
    }
    for (size_t q = 0; q < NG * kx; q++) {
        qg.p[q] = dmax(qg.p[q], 0.0);
        se.d[q] = cp * tg.p[q] + phig.p[q];
    }
    for (int k = 1; k <= kx; k++) {  // humidity.f90:17-28
        get_qsat(tg.p + NG * (k - 1), psg.d.data(), g.fsg[k], qsat.d.data() + NG * (k - 1), (int)NG);
        for (size_t q = 0; q < NG; q++) Q3(rh.v(), q, k) = Q3(qg, q, k) / Q3(qsat.v(), q, k);
    }
    G2 cbmf = s.g2(V_cbmf), precnv = s.g2(V_precnv), precls = s.g2(V_precls);
    get_convection_tendencies(psg, se, qg, qsat, iptop.data(), cbmf, precnv, tt_cnv, qt_cnv, g);
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) {
            Q3(tt_cnv.v(), q, k) = Q3(tt_cnv.v(), q, k) * rps.d[q] * g.grdscp[k];
            Q3(qt_cnv.v(), q, k) = Q3(qt_cnv.v(), q, k) * rps.d[q] * g.grdsig[k];
        }
    for (size_t q = 0; q < NG; q++) icnv[q] = kx - iptop[q];
    get_large_scale_condensation_tendencies(psg, qg, qsat, iptop.data(), precls, tt_lsc, qt_lsc, g);
    for (size_t q = 0; q < NG * kx; q++) {
        ttend.p[q] = ttend.p[q] + tt_cnv.d[q] + tt_lsc.d[q];
        qtend.p[q] = qtend.p[q] + qt_cnv.d[q] + qt_lsc.d[q];
    }
    G3 tt_rsw = s.g3(V_tt_rsw);
    if (s.compute_shortwave) {  // :151-169
        for (size_t q = 0; q < NG; q++)
            gse.d[q] = (Q3(se.v(), q, kx - 1) - Q3(se.v(), q, kx)) / (Q3(phig, q, kx - 1) - Q3(phig, q, kx));
        clouds(qg, rh, precnv, precls, iptop.data(), gse, s.g2(V_fmask_land), icltop.data(), cloudc, clstr,
               s.g2(V_qcloud_equiv));
        get_shortwave_rad_fluxes(s, psg, qg, icltop.data(), cloudc, clstr);
        for (int k = 1; k <= kx; k++)
            for (size_t q = 0; q < NG; q++) Q3(tt_rsw, q, k) = Q3(tt_rsw, q, k) * rps.d[q] * g.grdscp[k];
    }
    G2 slrd = s.g2(V_slrd);
    get_downward_longwave_rad_fluxes(s, tg, slrd, tt_rlw);
    get_surface_fluxes(s, psg, ug, vg, tg, qg, rh, phig, s.g2(V_sst_am), ts, tskin, u0, v0, t0);
    V3<double> slru = s.aux3(V_slru);
    get_upward_longwave_rad_fluxes(s, tg, ts, slrd, slru.slab(3), s.g2(V_slr), s.g2(V_olr), tt_rlw);
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < NG; q++) Q3(tt_rlw.v(), q, k) = Q3(tt_rlw.v(), q, k) * rps.d[q] * g.grdscp[k];
    for (size_t q = 0; q < NG * kx; q++) ttend.p[q] = ttend.p[q] + tt_rsw.p[q] + tt_rlw.d[q];

    get_vertical_diffusion_tend(se, rh, qg, qsat, phig, icnv.data(), ut_pbl, vt_pbl, tt_pbl, qt_pbl, g);
    V3<double> ustr = s.aux3(V_ustr), vstr = s.aux3(V_vstr), shf = s.aux3(V_shf), evap = s.aux3(V_evap);
    for (size_t q = 0; q < NG; q++) {
        Q3(ut_pbl.v(), q, kx) = Q3(ut_pbl.v(), q, kx) + Q3(ustr, q, 3) * rps.d[q] * g.grdsig[kx];
        Q3(vt_pbl.v(), q, kx) = Q3(vt_pbl.v(), q, kx) + Q3(vstr, q, 3) * rps.d[q] * g.grdsig[kx];
        Q3(tt_pbl.v(), q, kx) = Q3(tt_pbl.v(), q, kx) + Q3(shf, q, 3) * rps.d[q] * g.grdscp[kx];
        Q3(qt_pbl.v(), q, kx) = Q3(qt_pbl.v(), q, kx) + Q3(evap, q, 3) * rps.d[q] * g.grdsig[kx];
    }
    for (size_t q = 0; q < NG * kx; q++) {
        utend.p[q] = utend.p[q] + ut_pbl.d[q];
        vtend.p[q] = vtend.p[q] + vt_pbl.d[q];
        ttend.p[q] = ttend.p[q] + tt_pbl.d[q];
        qtend.p[q] = qtend.p[q] + qt_pbl.d[q];
    }
    if (dbg)
        for (size_t q = 0; q < NG; q++) dbg[q] = iptop[q], dbg[q + NG] = icnv[q], dbg[q + 2 * NG] = icltop[q];
}

}  // namespace orc
