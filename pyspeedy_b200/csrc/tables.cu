// speedy-b200: host-side generation of the shared model tables (one copy for the whole ensemble; the reference
// keeps one per member).
//
// The tables must carry gfortran's mixed REAL(4)/REAL(8) evaluation of the reference (no -fdefault-real-8:
// speedy.f90/Makefile:15-22), so everything here is computed on the host with glibc (cosf/sqrtf for the
// single-precision sites, cos/sin/log/pow for the double ones) and uploaded; nothing is regenerated with CUDA math.
// References: geometry.f90:61-156, fftpack.f90:1-67, legendre.f90:38-112,224-307, spectral.f90:39-116,
// horizontal_diffusion.f90:50-110, implicit.f90:44-218, matrix_inversion.f90, geopotential.f90:16-31,
// longwave_radiation.f90:208-232.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "spdy.cuh"
#include "tables.h"

namespace spdy {

// correctly-rounded REAL(4) values of constant expressions that gfortran folds at compile time
static const double F_ATAN1 = 0x1.921fb6p-1, F_ASIN1 = 0x1.921fb6p+0, F_SQRT2 = 0x1.6a09e6p+0,
                    F_SQRT3 = 0x1.bb67aep+0, F_LOG099 = -0x1.49544p-7;

const double H_REARTH = FL(6.371e+6), H_OMEGA = FL(7.292e-05), H_GRAV = FL(9.81), H_P0 = FL(1.e+5),
             H_CP = FL(1004.0), H_AKAP = (double)(2.0f / 7.0f), H_RGAS = H_AKAP *H_CP, H_GAMMA = 6.0,
             H_HSCALE = FL(7.5), H_HSHUM = FL(2.5), H_THD = FL(2.4), H_THDD = FL(2.4), H_THDS = 12.0,
             H_EPSLW = FL(0.05), H_DELT = 2400.0, H_ALPH = 0.5;

static inline int sidx(int m, int n) { return m + MX * n; }  // 0-based (m, n) -> linear, m fastest

// Gaussian weights, Newton iteration on P_48 (legendre.f90:224-257)
static void gauss_weights(double *w) {
    const int n = 2 * IY;
    double zprev = 2.0, dp = 0.0;
    for (int i = 0; i < IY; i++) {
        double z = cos(3.141592654 * ((double)(i + 1) - 0.25) / ((double)n + 0.5));
        while (fabs(z - zprev) > 2.220446049250313e-16) {
            double pa = 1.0, pb = 0.0;
            for (int j = 1; j <= n; j++) {
                double pc = pb;
                pb = pa;
                pa = ((2.0 * (double)j - 1.0) * z * pb - ((double)j - 1.0) * pc) / j;
            }
            dp = (double)n * (z * pa - pb) / (z * z - 1.0);
            zprev = z;
            z = zprev - pa / dp;
        }
        w[i] = 2.0 / ((1.0 - z * z) * (dp * dp));
    }
}

// LU inverse of an n x n Fortran-order matrix, same pivoting/ordering as matrix_inversion.f90:10-139
static void lu_inverse(double *a, double *y, int n) {
    std::vector<int> piv(n);
    std::vector<double> vv(n);
    auto A = [&](int i, int j) -> double & { return a[i + n * j]; };
    for (int i = 0; i < n; i++) {
        double big = 0.0;
        for (int j = 0; j < n; j++) big = fabs(A(i, j)) > big ? fabs(A(i, j)) : big;
        if (big == 0.0) {
            fprintf(stderr, "speedy-b200: singular matrix in the semi-implicit set-up\n");
            abort();
        }
        vv[i] = 1.0 / big;
    }
    int imax = 0;
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            double s = A(i, j);
            if (i > 0) {
                for (int k = 0; k < i; k++) s = s - A(i, k) * A(k, j);
                A(i, j) = s;
            }
        }
        double big = 0.0;
        for (int i = j; i < n; i++) {
            double s = A(i, j);
            if (j > 0) {
                for (int k = 0; k < j; k++) s = s - A(i, k) * A(k, j);
                A(i, j) = s;
            }
            double d = vv[i] * fabs(s);
            if (d >= big) {
                imax = i;
                big = d;
            }
        }
        if (j != imax) {
            for (int k = 0; k < n; k++) {
                double d = A(imax, k);
                A(imax, k) = A(j, k);
                A(j, k) = d;
            }
            vv[imax] = vv[j];
        }
        piv[j] = imax;
        if (j != n - 1) {
            if (A(j, j) == 0.0) A(j, j) = FL(1.0e-20);
            double d = 1.0 / A(j, j);
            for (int i = j + 1; i < n; i++) A(i, j) = A(i, j) * d;
        }
    }
    if (A(n - 1, n - 1) == 0.0) A(n - 1, n - 1) = FL(1.0e-20);
    for (int c = 0; c < n; c++) {
        double *b = y + n * c;
        for (int i = 0; i < n; i++) b[i] = (i == c) ? 1.0 : 0.0;
        int ii = -1;
        for (int i = 0; i < n; i++) {
            int ll = piv[i];
            double s = b[ll];
            b[ll] = b[i];
            if (ii >= 0) {
                for (int j = ii; j < i; j++) s = s - A(i, j) * b[j];
            } else if (s != 0.0) {
                ii = i;
            }
            b[i] = s;
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = b[i];
            for (int j = i + 1; j < n; j++) s = s - A(i, j) * b[j];
            b[i] = s / A(i, i);
        }
    }
}

static void build_impl(const ConstTables &C, const GlobTables &G, double dt, ImplTables &I) {
    for (int q = 0; q < NSPC; q++) {
        I.dmp1[q] = 1.0 / (1.0 + G.dmp[q] * dt);
        I.dmp1d[q] = 1.0 / (1.0 + G.dmpd[q] * dt);
        I.dmp1s[q] = 1.0 / (1.0 + G.dmps[q] * dt);
    }
    const double xi = dt * H_ALPH;
    const double xxi = xi / (H_REARTH * H_REARTH);
    for (int k = 0; k < KX; k++) I.dhsx[k] = xi * C.dhs[k];
    for (int n = 0; n < NX; n++)
        for (int m = 0; m < MX; m++) I.elz[sidx(m, n)] = (double)((float)(m + n) * (float)(m + n + 1)) * xxi;
    double xa[KX][KX] = {}, xb[KX][KX] = {}, xe[KX][KX], ya[KX][KX], xc[KX][KX], xd[KX][KX] = {}, dsum[KX];
    for (int k = 0; k < KX; k++)
        for (int k1 = 0; k1 < KX; k1++) ya[k][k1] = -H_AKAP * C.tref[k] * C.dhs[k1];
    for (int k = 1; k < KX; k++)
        xa[k][k - 1] = 0.5 * (H_AKAP * C.tref[k] / C.fsg[k] - (C.tref[k] - C.tref[k - 1]) / C.dhs[k]);
    for (int k = 0; k < KX - 1; k++)
        xa[k][k] = 0.5 * (H_AKAP * C.tref[k] / C.fsg[k] - (C.tref[k + 1] - C.tref[k]) / C.dhs[k]);
    dsum[0] = C.dhs[0];
    for (int k = 1; k < KX; k++) dsum[k] = dsum[k - 1] + C.dhs[k];
    for (int k = 0; k < KX - 1; k++)
        for (int k1 = 0; k1 < KX; k1++) {
            xb[k][k1] = C.dhs[k1] * dsum[k];
            if (k1 <= k) xb[k][k1] = xb[k][k1] - C.dhs[k1];
        }
    for (int k = 0; k < KX; k++)
        for (int k1 = 0; k1 < KX; k1++) {
            xc[k][k1] = ya[k][k1];
            for (int k2 = 0; k2 < KX - 1; k2++) xc[k][k1] = xc[k][k1] + xa[k][k2] * xb[k2][k1];
        }
    for (int k = 0; k < KX; k++)
        for (int k1 = k + 1; k1 < KX; k1++) xd[k][k1] = H_RGAS * log(C.hsg[k1 + 1] / C.hsg[k1]);
    for (int k = 0; k < KX; k++) xd[k][k] = H_RGAS * log(C.hsg[k + 1] / C.fsg[k]);
    for (int k = 0; k < KX; k++)
        for (int k1 = 0; k1 < KX; k1++) {
            xe[k][k1] = 0.0;
            for (int k2 = 0; k2 < KX; k2++) xe[k][k1] = xe[k][k1] + xd[k][k2] * xc[k2][k1];
        }
    double xf[KX * KX];
    for (int l = 1; l <= MX + NX + 1; l++) {
        const double xxx = (double)((float)l * (float)(l + 1)) / (H_REARTH * H_REARTH);
        for (int k = 0; k < KX; k++)
            for (int k1 = 0; k1 < KX; k1++)
                xf[k + KX * k1] = xi * xi * xxx * (H_RGAS * C.tref[k] * C.dhs[k1] - xe[k][k1]);
        for (int k = 0; k < KX; k++) xf[k + KX * k] = xf[k + KX * k] + 1.0;
        lu_inverse(xf, &I.xj[KX * KX * (l - 1)], KX);
    }
    for (int k = 0; k < KX; k++)
        for (int k1 = 0; k1 < KX; k1++) {
            I.xc[k + KX * k1] = xc[k][k1] * xi;
            I.xd[k + KX * k1] = xd[k][k1];
        }
}

void build_tables(ConstTables &C, GlobTables &G) {
    memset(&C, 0, sizeof(C));
    memset(&G, 0, sizeof(G));
    // ---- vertical grid (geometry.f90:89-102,133-153)
    const double hs[KX + 1] = {FL(0.000), FL(0.050), FL(0.140), FL(0.260), FL(0.420), FL(0.600), FL(0.770), FL(0.900), FL(1.000)};
    for (int k = 0; k <= KX; k++) C.hsg[k] = hs[k], C.sigh[k] = hs[k];
    for (int k = 0; k < KX; k++) {
        C.dhs[k] = C.hsg[k + 1] - C.hsg[k];
        C.fsg[k] = 0.5 * (C.hsg[k + 1] + C.hsg[k]);
        C.dhsr[k] = 0.5 / C.dhs[k];
        C.fsgr[k] = H_AKAP / (2.0 * C.fsg[k]);
        C.sigl[k] = log(C.fsg[k]);
        C.grdsig[k] = H_GRAV / (C.dhs[k] * H_P0);
        C.grdscp[k] = C.grdsig[k] / H_CP;
    }
    for (int k = 0; k < KX - 1; k++) {
        C.wvi[k][0] = 1.0 / (C.sigl[k + 1] - C.sigl[k]);
        C.wvi[k][1] = (log(C.sigh[k + 1]) - C.sigl[k]) * C.wvi[k][0];
    }
    C.wvi[KX - 1][0] = 0.0;
    C.wvi[KX - 1][1] = (F_LOG099 - C.sigl[KX - 1]) * C.wvi[KX - 2][0];
    // ---- column-physics constants (same operation order as the reference's per-call evaluation)
    {
        const double ENTMAX = FL(0.5), TRCNV = FL(6.0), EPSLW = FL(0.05);
        double entr[KX] = {0}, sentr = 0.0;
        for (int k = 1; k < KX - 1; k++) {
            const double ee = fmax(0.0, C.fsg[k] - 0.5);
            entr[k] = ee * ee;
            sentr = sentr + entr[k];
        }
        sentr = ENTMAX / sentr;
        for (int k = 0; k < KX; k++) C.ph_entrs[k] = entr[k] * sentr;
        C.ph_fm0 = H_P0 * C.dhs[7] / (H_GRAV * TRCNV * FL(3600.0));
        C.ph_eps1 = EPSLW / (C.dhs[0] + C.dhs[1]);
        const double trshc = FL(6.0), trvdi = FL(24.0), trvds = FL(6.0), rhgrad = FL(0.5);
        const double cshc = C.dhs[7] / FL(3600.0);
        const double cvdi = (C.sigh[7] - C.sigh[1]) / (double)(6.0f * 3600.0f);
        C.ph_fshcq = cshc / trshc, C.ph_fshcse = cshc / (trshc * H_CP);
        C.ph_fvdiq = cvdi / trvdi, C.ph_fvdise = cvdi / (trvds * H_CP);
        for (int k = 0; k < KX; k++) {
            C.ph_rdhs[k] = 1.0 / C.dhs[k];
            C.ph_fvdiq2[k] = C.ph_fvdiq * C.sigh[k];
            if (k >= 1) C.ph_drh0[k] = rhgrad * (C.fsg[k] - C.fsg[k - 1]);
            if (k < KX - 1) C.ph_r1sig[k] = 1.0 / (1.0 - C.sigh[k + 1]);
        }
        C.ph_rt1s = 1.0 / (H_RGAS * FL(288.0) * C.sigl[KX - 1]);
    }
    // ---- latitudes (geometry.f90:107-131): REAL(4) argument and cosf
    double sia_half[IY], coa_half[IY];
    for (int j = 0; j < IY; j++) {
        const float arg = 3.141592654f * ((float)(j + 1) - 0.25f) / ((float)IL + 0.5f);
        sia_half[j] = (double)cosf(arg);
        coa_half[j] = sqrt(1.0 - sia_half[j] * sia_half[j]);
        const int jn = IL - 1 - j;
        C.sia[j] = -sia_half[j], C.sia[jn] = sia_half[j];
        C.coa[j] = C.coa[jn] = coa_half[j];
        C.radang[j] = -asin(sia_half[j]), C.radang[jn] = asin(sia_half[j]);
        C.cosgr[j] = C.cosgr[jn] = 1.0 / coa_half[j];
        C.cosgr2[j] = C.cosgr2[jn] = 1.0 / (coa_half[j] * coa_half[j]);
        C.ph_sqcoa[j] = C.ph_sqcoa[jn] = sqrt(coa_half[j]);
    }
    for (int j = 0; j < IL; j++) C.coriol[j] = 2.0 * H_OMEGA * C.sia[j];
    // ---- FFT twiddles (fftpack.f90:39-66) for the factor sequence 2,4,4,3
    {
        const int fac[4] = {2, 4, 4, 3};
        const double tpi = 8.0 * F_ATAN1, argh = tpi / IX;
        int is = 0, l1 = 1;
        for (int p = 0; p < 3; p++) {
            const int ip = fac[p], l2 = l1 * ip, ido = IX / l2;
            int ld = 0;
            for (int jj = 1; jj < ip; jj++) {
                ld += l1;
                const double argld = ld * argh;
                double fi = 0.0;
                int i = is;
                for (int ii = 3; ii <= ido; ii += 2) {
                    i += 2;
                    fi += 1.0;
                    C.wa[i - 2] = cos(fi * argld);
                    C.wa[i - 1] = sin(fi * argld);
                }
                is += ido;
            }
            l1 = l2;
        }
        C.fc[0] = 0.5 * F_SQRT3;          // taui  (fftpack.f90:269,787)
        C.fc[1] = F_SQRT2;                // sqrt2 (:341)
        C.fc[2] = 0.5 * F_SQRT2;          // hsqt2 (:857)
        C.fc[3] = (double)(1.0f / 96.0f); // fourier.f90:113
    }
    // ---- Legendre (legendre.f90:64-108, 260-307)
    gauss_weights(C.wt);
    {
        std::vector<double> eps((MX + 1) * (NX + 1)), reps((MX + 1) * (NX + 1));
        auto E = [&](int m, int n) -> double & { return eps[m + (MX + 1) * n]; };
        auto R = [&](int m, int n) -> double & { return reps[m + (MX + 1) * n]; };
        for (int m = 0; m <= MX; m++)
            for (int n = 0; n <= NX; n++) {
                const float fm = (float)m, fl = (float)(n + m);
                const double emm2 = (double)(fm * fm), ell2 = (double)(fl * fl);
                E(m, n) = (n == NX || (n == 0 && m == 0)) ? 0.0 : sqrt((ell2 - emm2) / (4.0 * ell2 - 1.0));
                R(m, n) = E(m, n) > 0.0 ? 1.0 / E(m, n) : 0.0;
            }
        std::vector<double> alp((MX + 1) * NX);
        auto P = [&](int m, int n) -> double & { return alp[m + (MX + 1) * n]; };
        for (int j = 0; j < IY; j++) {
            const double y = coa_half[j], x = sia_half[j];
            P(0, 0) = 0x1.6a09e6p-1;  // sqrt(0.5) REAL(4)
            for (int m = 1; m <= MX; m++) {
                const double consq = (double)sqrtf(0.5f * (2.0f * (float)m + 1.0f) / (float)m);
                P(m, 0) = consq * y * P(m - 1, 0);
            }
            for (int m = 0; m <= MX; m++) P(m, 1) = (x * P(m, 0)) * R(m, 1);
            for (int n = 2; n < NX; n++)
                for (int m = 0; m <= MX; m++) P(m, n) = (x * P(m, n - 1) - E(m, n - 1) * P(m, n - 2)) * R(m, n);
            for (int n = 0; n < NX; n++)
                for (int m = 0; m < MX; m++) {
                    double v = P(m, n);
                    if (fabs(v) <= FL(1.e-30)) v = 0.0;
                    G.cpol[(m * NX + n) * IY + j] = v;
                }
        }
        // ---- parity-pure DMMA fragments of the fused spec->grid kernel (fused_mma3.cu): M = 8 latitudes
        {
            int koff = 0;
            for (int m = 0; m < MX; m++) {
                const int nmax = 31 - m, cnt[2] = {(32 - m + 1) / 2, (32 - m) / 2};
                const int ks[2] = {(cnt[0] + 3) / 4, (cnt[1] + 3) / 4};
                for (int jo = 0; jo < IY / 8; jo++)
                    for (int par = 0; par < 2; par++)
                        for (int s = 0; s < ks[par]; s++)
                            for (int L = 0; L < 32; L++) {
                                const int n = par + 2 * (4 * s + (L & 3)), j = 8 * jo + (L >> 2);
                                G.pq_inv2[((size_t)jo * PQ2_KTOT + koff + (par ? ks[0] : 0) + s) * 32 + L] =
                                    (n <= nmax) ? G.cpol[(m * NX + n) * IY + j] : 0.0;
                            }
                koff += ks[0] + ks[1];
            }
            if (koff != PQ2_KTOT) abort();
        }
        // ---- parity-pure DMMA fragments of the fused grid->spec kernel (fused_mma2.cu): the N/S fold
        //      (legendre.f90:196-203) is done on the Fourier rows, so a tile needs one k-slice instead of two
        {
            int toff = 0;
            for (int m = 0; m < MX; m++) {
                const int nmax = (30 < 31 - m) ? 30 : 31 - m;
                const int cnt[2] = {nmax / 2 + 1, (nmax + 1) / 2};
                const int nt[2] = {(cnt[0] + 7) / 8, (cnt[1] + 7) / 8};
                for (int jq = 0; jq < IY / 4; jq++)
                    for (int par = 0; par < 2; par++)
                        for (int i = 0; i < nt[par]; i++)
                            for (int L = 0; L < 32; L++) {
                                const int n = par + 2 * (8 * i + (L >> 2)), j = 4 * jq + (L & 3);
                                G.pq_dir2[((size_t)jq * PD2_TTOT + toff + (par ? nt[0] : 0) + i) * 32 + L] =
                                    (n <= nmax) ? C.wt[j] * G.cpol[(m * NX + n) * IY + j] : 0.0;
                            }
                toff += nt[0] + nt[1];
            }
            if (toff != PD2_TTOT) abort();
        }
        // ---- SPPT amplitude spectrum (sppt.f90:27-37,87-95); filled after el2 below
        // ---- spectral operator tables (spectral.f90:68-110)
        const double re2 = H_REARTH * H_REARTH;
        for (int n = 0; n < NX; n++)
            for (int m = 0; m < MX; m++) {
                const int l = m + n, q = sidx(m, n);
                G.el2[q] = (double)(float)(l * (l + 1)) / re2;
                G.elm2[q] = (l == 0) ? 0.0 : 1.0 / G.el2[q];
                G.trfilt[q] = (l <= NTRUNC) ? 1.0 : 0.0;
                const double el1 = (double)(float)l;
                if (n == 0) {
                    G.gradx[m] = (double)(float)m / H_REARTH;
                    G.uvdx[q] = -H_REARTH / (double)(float)(m + 1);
                } else {
                    G.uvdx[q] = -H_REARTH * (double)(float)m / (el1 * (el1 + 1));
                    G.gradym[q] = (el1 - 1.0) * E(m, n) / H_REARTH;
                    G.uvdym[q] = -H_REARTH * E(m, n) / el1;
                    G.vddym[q] = (el1 + 1) * E(m, n) / H_REARTH;
                }
                G.gradyp[q] = (el1 + 2.0) * E(m, n + 1) / H_REARTH;
                G.uvdyp[q] = -H_REARTH * E(m, n + 1) / (el1 + 1.0);
                G.vddyp[q] = el1 * E(m, n + 1) / H_REARTH;
            }
        // ---- SPPT amplitude spectrum and AR(1) constants (sppt.f90:27-37,87-95)
        {
            const double time_decorr = 6.0, len_decorr = 500000.0, stddev = FL(0.33);
            const double phi = exp(-(24 / (double)NSTEPS) / time_decorr);
            double f0 = 0.0;
            for (int n = 1; n <= NTRUNC; n++)
                f0 = f0 + (2 * n + 1) * exp(-0.5 * ((len_decorr / H_REARTH) * (len_decorr / H_REARTH)) * n * (n + 1));
            f0 = sqrt((stddev * stddev * (1 - phi * phi)) / (2 * f0));
            for (int q = 0; q < NSPC; q++) G.sppt_sigma[q] = f0 * exp(-0.25 * (len_decorr * len_decorr) * G.el2[q]);
            G.sppt_phi = phi, G.sppt_first_fac = pow(1 - phi * phi, -0.5);
        }
    }
    // ---- horizontal diffusion (horizontal_diffusion.f90:76-108)
    {
        const double hdiff = 1.0 / (H_THD * FL(3600.)), hdifd = 1.0 / (H_THDD * FL(3600.)), hdifs = 1.0 / (H_THDS * FL(3600.));
        const double rlap = (double)(1.0f / (float)(NTRUNC * (NTRUNC + 1)));
        for (int n = 0; n < NX; n++)
            for (int m = 0; m < MX; m++) {
                const double twn = (double)(float)(m + n);
                const double elap = twn * (twn + 1.0) * rlap;
                const double e2 = elap * elap, elapn = e2 * e2;
                G.dmp[sidx(m, n)] = hdiff * elapn;
                G.dmpd[sidx(m, n)] = hdifd * elapn;
                G.dmps[sidx(m, n)] = hdifs * elap;
            }
        const double rgam = H_RGAS * H_GAMMA / (FL(1000.) * H_GRAV), qexp = H_HSCALE / H_HSHUM;
        for (int k = 1; k < KX; k++) {
            C.tcorv[k] = pow(C.fsg[k], rgam);
            if (k > 1) C.qcorv[k] = pow(C.fsg[k], qexp);
        }
        for (int k = 0; k < KX; k++) {  // implicit.f90:70-78
            const double f = C.fsg[k] > FL(0.2) ? C.fsg[k] : FL(0.2);
            C.tref[k] = FL(288.) * pow(f, rgam);
            C.tref2[k] = H_AKAP * C.tref[k];
            C.tref3[k] = C.fsgr[k] * C.tref[k];
        }
    }
    // ---- geopotential (geopotential.f90:26-29,72-76)
    for (int k = 0; k < KX; k++) {
        C.xgeop1[k] = H_RGAS * log(C.hsg[k + 1] / C.fsg[k]);
        if (k != KX - 1) C.xgeop2[k + 1] = H_RGAS * log(C.fsg[k + 1] / C.hsg[k + 1]);
    }
    for (int k = 1; k < KX - 1; k++)
        C.geocorf[k] = C.xgeop1[k] * 0.5 * log(C.hsg[k + 1] / C.fsg[k]) / log(C.fsg[k + 1] / C.fsg[k - 1]);
    // ---- long-wave band fractions (longwave_radiation.f90:208-232): polynomials in REAL(4)
    {
        const double eps1 = 1.0 - H_EPSLW;
        auto F = [&](int t, int b) -> double & { return G.fband[(t - 100) + 301 * b]; };
        for (int t = 200; t <= 320; t++) {
            F(t, 1) = (double)(0.148f - 3.0e-6f * (float)((t - 247) * (t - 247))) * eps1;
            F(t, 2) = (double)(0.356f - 5.2e-6f * (float)((t - 282) * (t - 282))) * eps1;
            F(t, 3) = (double)(0.314f + 1.0e-5f * (float)((t - 315) * (t - 315))) * eps1;
            F(t, 0) = eps1 - (F(t, 1) + F(t, 2) + F(t, 3));
        }
        for (int b = 0; b < 4; b++) {
            for (int t = 100; t < 200; t++) F(t, b) = F(200, b);
            for (int t = 321; t <= 400; t++) F(t, b) = F(320, b);
        }
    }
    // ---- semi-implicit tables for the three time steps used (time_stepping.f90:13-27)
    build_impl(C, G, 0.5 * H_DELT, G.impl[0]);
    build_impl(C, G, H_DELT, G.impl[1]);
    build_impl(C, G, 2.0 * H_DELT, G.impl[2]);
    memcpy(C.xc2, G.impl[2].xc, sizeof(C.xc2));
    memcpy(C.xd2, G.impl[2].xd, sizeof(C.xd2));
    memcpy(C.xj2, G.impl[2].xj, sizeof(C.xj2));
    memcpy(C.dhsx2, G.impl[2].dhsx, sizeof(C.dhsx2));
}

}  // namespace spdy
