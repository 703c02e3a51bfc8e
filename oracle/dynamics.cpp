// ORACLE (test infrastructure) -- geopotential, grid-point / spectral tendencies, semi-implicit scheme,
// horizontal diffusion tables, time stepping and the diagnostics check.
// Follows geopotential.f90, tendencies.f90, implicit.f90, matrix_inversion.f90, horizontal_diffusion.f90,
// time_stepping.f90 and diagnostics.f90 of the reference.
#include <cstdio>
#include <cstdlib>

#include "speedy_oracle.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------------
// geopotential.f90:16-31
void initialize_geopotential(State &s) {
    const Geometry &g = s.geo;
    double *xgeop1 = s.p(V_xgeop1) - 1, *xgeop2 = s.p(V_xgeop2) - 1;
    for (int k = 1; k <= kx; k++) {
        xgeop1[k] = rgas * log(g.hsg[k + 1] / g.fsg[k]);
        if (k != kx) xgeop2[k + 1] = rgas * log(g.fsg[k + 1] / g.hsg[k + 1]);
    }
}

// geopotential.f90:36-77
void set_geopotential(State &s, int time_level) {
    const Geometry &g = s.geo;
    S3 t = s.s4lev(V_t, time_level), phi = s.s3(V_phi);
    S2 phis = s.s2(V_phis);
    const double *xgeop1 = s.p(V_xgeop1) - 1, *xgeop2 = s.p(V_xgeop2) - 1;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) phi(m, n, kx) = phis(m, n) + xgeop1[kx] * t(m, n, kx);
    for (int k = kx - 1; k >= 1; k--)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++)
                phi(m, n, k) = phi(m, n, k + 1) + xgeop2[k + 1] * t(m, n, k + 1) + xgeop1[k] * t(m, n, k);
    for (int k = 2; k <= kx - 1; k++) {
        double corf = xgeop1[k] * 0.5 * log(g.hsg[k + 1] / g.fsg[k]) / log(g.fsg[k + 1] / g.fsg[k - 1]);
        for (int n = 1; n <= nx; n++) phi(1, n, k) = phi(1, n, k) + corf * (t(1, n, k + 1) - t(1, n, k - 1));
    }
}

// ---------------------------------------------------------------------------------------------------
// tendencies.f90:51-276
void get_grid_point_tendencies(State &s, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt, int j1, int j2) {
    const Spectral &sp = s.spec;
    const Implicit &im = s.imp;
    const Geometry &g = s.geo;
    Spec2 dumc1, dumc2;
    Grid3 utend, vtend, ttend, trtend, ug, vg, tg, vorg, divg, tgg, puv, trg;
    Grid2 px, py, umean, vmean, dmean;
    Grid3 sigdt(kx + 1), temp(kx + 1), sigm(kx + 1);
    const size_t ng = (size_t)ix * il;

    S3 vor = s.s4lev(V_vor, j2), div = s.s4lev(V_div, j2), t = s.s4lev(V_t, j2), tr = s.s4lev(V_tr, j2);
    for (int k = 1; k <= kx; k++) {  // :109-130
        sp.spec2grid(vor.slab(k), vorg.slab(k), 1);
        sp.spec2grid(div.slab(k), divg.slab(k), 1);
        sp.spec2grid(t.slab(k), tg.slab(k), 1);
        sp.spec2grid(tr.slab(k), trg.slab(k), 1);
        sp.vort2vel(vor.slab(k), div.slab(k), dumc1, dumc2);
        sp.spec2grid(dumc2, vg.slab(k), 2);
        sp.spec2grid(dumc1, ug.slab(k), 2);
        for (int j = 1; j <= il; j++)
            for (int i = 1; i <= ix; i++) vorg(i, j, k) = vorg(i, j, k) + g.coriol[j];
    }
    for (int k = 1; k <= kx; k++)  // :136-140
        for (size_t q = 0; q < ng; q++) {
            umean.d[q] = umean.d[q] + ug.d[q + ng * (k - 1)] * g.dhs[k];
            vmean.d[q] = vmean.d[q] + vg.d[q + ng * (k - 1)] * g.dhs[k];
            dmean.d[q] = dmean.d[q] + divg.d[q + ng * (k - 1)] * g.dhs[k];
        }
    sp.gradient(s.s3lev(V_ps, j2), dumc1, dumc2);  // :144-146
    sp.spec2grid(dumc1, px, 2);
    sp.spec2grid(dumc2, py, 2);
    {
        Grid2 tmp;
        for (size_t q = 0; q < ng; q++) tmp.d[q] = -umean.d[q] * px.d[q] - vmean.d[q] * py.d[q];
        sp.grid2spec(tmp, psdt);
        psdt(1, 1) = cplx{0.0, 0.0};
    }
    // :151-165  (sigdt, sigm level 1 and kx+1 are zero from construction)
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            puv.d[q + ng * (k - 1)] = (ug.d[q + ng * (k - 1)] - umean.d[q]) * px.d[q] +
                                      (vg.d[q + ng * (k - 1)] - vmean.d[q]) * py.d[q];
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++) {
            sigdt.d[q + ng * k] = sigdt.d[q + ng * (k - 1)] -
                                  g.dhs[k] * (puv.d[q + ng * (k - 1)] + divg.d[q + ng * (k - 1)] - dmean.d[q]);
            sigm.d[q + ng * k] = sigm.d[q + ng * (k - 1)] - g.dhs[k] * puv.d[q + ng * (k - 1)];
        }
    // NB the reference sets sigdt(:,:,kx+1) = 0 BEFORE the loop, which then overwrites it (:152,161-164)
    for (int k = 1; k <= kx; k++)  // :169-171
        for (size_t q = 0; q < ng; q++) tgg.d[q + ng * (k - 1)] = tg.d[q + ng * (k - 1)] - im.tref[k];

#define K_(a, k) a.d[q + ng * ((k)-1)]
    // zonal wind tendency :174-184  (temp level 1 and kx+1 stay zero)
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < ng; q++) K_(temp, k) = K_(sigdt, k) * (K_(ug, k) - K_(ug, k - 1));
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            K_(utend, k) = K_(vg, k) * K_(vorg, k) - K_(tgg, k) * rgas * px.d[q] -
                           (K_(temp, k + 1) + K_(temp, k)) * g.dhsr[k];
    // meridional wind tendency :187-194
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < ng; q++) K_(temp, k) = K_(sigdt, k) * (K_(vg, k) - K_(vg, k - 1));
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            K_(vtend, k) = -K_(ug, k) * K_(vorg, k) - K_(tgg, k) * rgas * py.d[q] -
                           (K_(temp, k + 1) + K_(temp, k)) * g.dhsr[k];
    // temperature tendency :197-209
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            K_(temp, k) = K_(sigdt, k) * (K_(tgg, k) - K_(tgg, k - 1)) + K_(sigm, k) * (im.tref[k] - im.tref[k - 1]);
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            K_(ttend, k) = K_(tgg, k) * K_(divg, k) - (K_(temp, k + 1) + K_(temp, k)) * g.dhsr[k] +
                           g.fsgr[k] * K_(tgg, k) * (K_(sigdt, k + 1) + K_(sigdt, k)) +
                           im.tref3[k] * (K_(sigm, k + 1) + K_(sigm, k)) +
                           akap * (K_(tg, k) * K_(puv, k) - K_(tgg, k) * dmean.d[q]);
    // tracer tendency :212-224
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < ng; q++) K_(temp, k) = K_(sigdt, k) * (K_(trg, k) - K_(trg, k - 1));
    for (int k = 2; k <= 3; k++)
        for (size_t q = 0; q < ng; q++) K_(temp, k) = 0.0;
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ng; q++)
            K_(trtend, k) = K_(trg, k) * K_(divg, k) - (K_(temp, k + 1) + K_(temp, k)) * g.dhsr[k];

    // physics :229-232
    set_geopotential(s, j1);
    get_physical_tendencies(s, j1, utend, vtend, ttend, trtend);

    // back to spectral space :238-268
    Grid2 a, b;
    Spec2 sa, lap;
    for (int k = 1; k <= kx; k++) {
        sp.grid_vel2vort(utend.slab(k), vtend.slab(k), vordt.slab(k), divdt.slab(k), 2);
        for (size_t q = 0; q < ng; q++) a.d[q] = 0.5 * (K_(ug, k) * K_(ug, k) + K_(vg, k) * K_(vg, k));
        sp.grid2spec(a, sa);
        sp.laplacian(sa, lap);
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) divdt(m, n, k) = divdt(m, n, k) - lap(m, n);

        for (size_t q = 0; q < ng; q++) {
            a.d[q] = -K_(ug, k) * K_(tgg, k);
            b.d[q] = -K_(vg, k) * K_(tgg, k);
        }
        sp.grid_vel2vort(a, b, dumc1, tdt.slab(k), 2);
        sp.grid2spec(ttend.slab(k), sa);
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) tdt(m, n, k) = tdt(m, n, k) + sa(m, n);

        for (size_t q = 0; q < ng; q++) {
            a.d[q] = -K_(ug, k) * K_(trg, k);
            b.d[q] = -K_(vg, k) * K_(trg, k);
        }
        sp.grid_vel2vort(a, b, dumc1, trdt.slab(k), 2);
        sp.grid2spec(trtend.slab(k), sa);
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) trdt(m, n, k) = trdt(m, n, k) + sa(m, n);
    }
#undef K_
}

// tendencies.f90:283-352
void get_spectral_tendencies(State &s, S3 divdt, S3 tdt, S2 psdt, int j2) {
    const Spectral &sp = s.spec;
    const Implicit &im = s.imp;
    const Geometry &g = s.geo;
    Spec3 dumk(kx + 1), sigdtc(kx + 1);
    Spec2 dmeanc;
    S3 div = s.s4lev(V_div, j2);
    S2 ps = s.s3lev(V_ps, j2);
    const size_t ns = (size_t)mx * nx;
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ns; q++) dmeanc.d[q] = dmeanc.d[q] + div.p[q + ns * (k - 1)] * g.dhs[k];
    for (size_t q = 0; q < ns; q++) psdt.p[q] = psdt.p[q] - dmeanc.d[q];
    psdt(1, 1) = cplx{0.0, 0.0};
    for (int k = 1; k <= kx - 1; k++)
        for (size_t q = 0; q < ns; q++)
            sigdtc.d[q + ns * k] = sigdtc.d[q + ns * (k - 1)] - g.dhs[k] * (div.p[q + ns * (k - 1)] - dmeanc.d[q]);
    for (int k = 2; k <= kx; k++)
        for (size_t q = 0; q < ns; q++) dumk.d[q + ns * (k - 1)] = sigdtc.d[q + ns * (k - 1)] * (im.tref[k] - im.tref[k - 1]);
    for (int k = 1; k <= kx; k++)
        for (size_t q = 0; q < ns; q++)
            tdt.p[q + ns * (k - 1)] = tdt.p[q + ns * (k - 1)] -
                                      (dumk.d[q + ns * k] + dumk.d[q + ns * (k - 1)]) * g.dhsr[k] +
                                      im.tref3[k] * (sigdtc.d[q + ns * k] + sigdtc.d[q + ns * (k - 1)]) -
                                      im.tref2[k] * dmeanc.d[q];
    set_geopotential(s, j2);
    S3 phi = s.s3(V_phi);
    Spec2 tmp, lap;
    for (int k = 1; k <= kx; k++) {
        for (size_t q = 0; q < ns; q++) tmp.d[q] = phi.p[q + ns * (k - 1)] + (rgas * im.tref[k]) * ps.p[q];
        sp.laplacian(tmp, lap);
        for (size_t q = 0; q < ns; q++) divdt.p[q + ns * (k - 1)] = divdt.p[q + ns * (k - 1)] - lap.d[q];
    }
}

// tendencies.f90:11-39
void get_tendencies(State &s, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt, int j2) {
    get_grid_point_tendencies(s, vordt, divdt, tdt, psdt, trdt, 1, j2);
    if (alph < 0.5) {
        get_spectral_tendencies(s, divdt, tdt, psdt, j2);
    } else {
        get_spectral_tendencies(s, divdt, tdt, psdt, 1);
        s.imp.implicit_terms(divdt, tdt, psdt);
    }
}

// ---------------------------------------------------------------------------------------------------
// matrix_inversion.f90:10-139 (Numerical-Recipes style LU; a and y are (n,n) Fortran order)
static void ludcmp(double *a, int n, int *indx, double &d) {
    const double tiny = FL(1.0e-20);
    double vv[100];
    auto A = [&](int i, int j) -> double & { return a[(i - 1) + (size_t)n * (j - 1)]; };
    d = 1.0;
    int imax = 0;
    for (int i = 1; i <= n; i++) {
        double aamax = 0.0;
        for (int j = 1; j <= n; j++)
            if (fabs(A(i, j)) > aamax) aamax = fabs(A(i, j));
        if (aamax == 0.0) {
            fprintf(stderr, "Error during the LU decomposition. The input matrix is singular.\n");
            abort();
        }
        vv[i] = 1.0 / aamax;
    }
    for (int j = 1; j <= n; j++) {
        if (j > 1)
            for (int i = 1; i <= j - 1; i++) {
                double sum = A(i, j);
                if (i > 1) {
                    for (int k = 1; k <= i - 1; k++) sum = sum - A(i, k) * A(k, j);
                    A(i, j) = sum;
                }
            }
        double aamax = 0.0;
        for (int i = j; i <= n; i++) {
            double sum = A(i, j);
            if (j > 1) {
                for (int k = 1; k <= j - 1; k++) sum = sum - A(i, k) * A(k, j);
                A(i, j) = sum;
            }
            double dum = vv[i] * fabs(sum);
            if (dum >= aamax) {
                imax = i;
                aamax = dum;
            }
        }
        if (j != imax) {
            for (int k = 1; k <= n; k++) {
                double dum = A(imax, k);
                A(imax, k) = A(j, k);
                A(j, k) = dum;
            }
            d = -d;
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (j != n) {
            if (A(j, j) == 0.0) A(j, j) = tiny;
            double dum = 1.0 / A(j, j);
            for (int i = j + 1; i <= n; i++) A(i, j) = A(i, j) * dum;
        }
    }
    if (A(n, n) == 0.0) A(n, n) = tiny;
}

static void lubksb(const double *a, int n, const int *indx, double *b /*1-based*/) {
    auto A = [&](int i, int j) { return a[(i - 1) + (size_t)n * (j - 1)]; };
    int ii = 0;
    for (int i = 1; i <= n; i++) {
        int ll = indx[i];
        double sum = b[ll];
        b[ll] = b[i];
        if (ii != 0) {
            for (int j = ii; j <= i - 1; j++) sum = sum - A(i, j) * b[j];
        } else if (sum != 0.0) {
            ii = i;
        }
        b[i] = sum;
    }
    for (int i = n; i >= 1; i--) {
        double sum = b[i];
        if (i < n)
            for (int j = i + 1; j <= n; j++) sum = sum - A(i, j) * b[j];
        b[i] = sum / A(i, i);
    }
}

void inv8(double *a, double *y, int n) {
    int indx[101];
    double d;
    for (int i = 0; i < n * n; i++) y[i] = 0.0;
    for (int i = 1; i <= n; i++) y[(i - 1) + (size_t)n * (i - 1)] = 1.0;
    ludcmp(a, n, indx, d);
    for (int i = 1; i <= n; i++) lubksb(a, n, indx, y + (size_t)n * (i - 1) - 1);
}

// ---------------------------------------------------------------------------------------------------
// horizontal_diffusion.f90:50-110 and implicit.f90:44-81
void Implicit::initialize(const Geometry *g) {
    geo = g;
    const size_t ns = (size_t)mx * nx;
    dmp.assign(ns, 0.0), dmpd.assign(ns, 0.0), dmps.assign(ns, 0.0);
    dmp1.assign(ns, 0.0), dmp1d.assign(ns, 0.0), dmp1s.assign(ns, 0.0);
    tcorh.assign(ns, cplx{0.0, 0.0}), qcorh.assign(ns, cplx{0.0, 0.0});
    xj.assign((size_t)kx * kx * (mx + nx + 1), 0.0);
    elz.assign(ns, 0.0);

    const double hdiff = 1.0 / (thd * FL(3600.)), hdifd = 1.0 / (thdd * FL(3600.)), hdifs = 1.0 / (thds * FL(3600.));
    const double rlap = (double)(1.0f / (float)(trunc_ * (trunc_ + 1)));  // :82 REAL(4) division
    for (int j = 1; j <= nx; j++)
        for (int k = 1; k <= mx; k++) {
            double twn = (double)(float)(k + j - 2);
            double elap = (twn * (twn + 1.0) * rlap);
            double e2 = elap * elap;
            double elapn = e2 * e2;  // elap**4, integer power
            dmp[Spectral::i2(k, j)] = hdiff * elapn;
            dmpd[Spectral::i2(k, j)] = hdifd * elapn;
            dmps[Spectral::i2(k, j)] = hdifs * elap;
        }
    double rgam = rgas * gamma_ / (FL(1000.) * grav);
    double qexp = hscale / hshum;
    tcorv[1] = 0.0;
    qcorv[1] = 0.0;
    qcorv[2] = 0.0;
    for (int k = 2; k <= kx; k++) {
        tcorv[k] = pow(g->fsg[k], rgam);
        if (k > 2) qcorv[k] = pow(g->fsg[k], qexp);
    }
    // implicit.f90:70-78
    for (int k = 1; k <= kx; k++) {
        tref[k] = FL(288.) * pow(fmax(FL(0.2), g->fsg[k]), rgam);
        tref2[k] = akap * tref[k];
        tref3[k] = g->fsgr[k] * tref[k];
    }
}

// implicit.f90:83-218
void Implicit::set_time_step(double dt) {
    const Geometry &g = *geo;
    double dsum[kx + 1], xa[kx + 1][kx + 1], xb[kx + 1][kx + 1], xe[kx + 1][kx + 1], ya[kx + 1][kx + 1];
    std::vector<double> xf((size_t)kx * kx * (mx + nx + 1));
    auto XF = [&](int k, int k1, int l) -> double & { return xf[(k - 1) + (size_t)kx * ((k1 - 1) + (size_t)kx * (l - 1))]; };
    for (int m = 1; m <= mx; m++)
        for (int n = 1; n <= nx; n++) {
            size_t q = Spectral::i2(m, n);
            dmp1[q] = 1.0 / (1.0 + dmp[q] * dt);
            dmp1d[q] = 1.0 / (1.0 + dmpd[q] * dt);
            dmp1s[q] = 1.0 / (1.0 + dmps[q] * dt);
        }
    double xi = dt * alph;
    double xxi = xi / (rearth * rearth);
    for (int k = 1; k <= kx; k++) dhsx[k] = xi * g.dhs[k];
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) elz[Spectral::i2(m, n)] = (double)((float)(m + n - 2) * (float)(m + n - 1)) * xxi;
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) xa[k][k1] = 0.0, xb[k][k1] = 0.0;
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) ya[k][k1] = -akap * tref[k] * g.dhs[k1];
    for (int k = 2; k <= kx; k++)
        xa[k][k - 1] = 0.5 * (akap * tref[k] / g.fsg[k] - (tref[k] - tref[k - 1]) / g.dhs[k]);
    for (int k = 1; k <= kx - 1; k++)
        xa[k][k] = 0.5 * (akap * tref[k] / g.fsg[k] - (tref[k + 1] - tref[k]) / g.dhs[k]);
    dsum[1] = g.dhs[1];
    for (int k = 2; k <= kx; k++) dsum[k] = dsum[k - 1] + g.dhs[k];
    for (int k = 1; k <= kx - 1; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xb[k][k1] = g.dhs[k1] * dsum[k];
            if (k1 <= k) xb[k][k1] = xb[k][k1] - g.dhs[k1];
        }
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xc[k][k1] = ya[k][k1];
            for (int k2 = 1; k2 <= kx - 1; k2++) xc[k][k1] = xc[k][k1] + xa[k][k2] * xb[k2][k1];
        }
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) xd[k][k1] = 0.0;
    for (int k = 1; k <= kx; k++)
        for (int k1 = k + 1; k1 <= kx; k1++) xd[k][k1] = rgas * log(g.hsg[k1 + 1] / g.hsg[k1]);
    for (int k = 1; k <= kx; k++) xd[k][k] = rgas * log(g.hsg[k + 1] / g.fsg[k]);
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xe[k][k1] = 0.0;
            for (int k2 = 1; k2 <= kx; k2++) xe[k][k1] = xe[k][k1] + xd[k][k2] * xc[k2][k1];
        }
    for (int l = 1; l <= mx + nx + 1; l++) {
        double xxx = (double)((float)l * (float)(l + 1)) / (rearth * rearth);
        for (int k = 1; k <= kx; k++)
            for (int k1 = 1; k1 <= kx; k1++) XF(k, k1, l) = xi * xi * xxx * (rgas * tref[k] * g.dhs[k1] - xe[k][k1]);
        for (int k = 1; k <= kx; k++) XF(k, k, l) = XF(k, k, l) + 1.0;
    }
    for (int l = 1; l <= mx + nx + 1; l++) inv8(&XF(1, 1, l), &XJ(1, 1, l), kx);
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) xc[k][k1] = xc[k][k1] * xi;
}

// implicit.f90:234-289
void Implicit::implicit_terms(S3 divdt, S3 tdt, S2 psdt) {
    Spec3 ye, yf;
    for (int k1 = 1; k1 <= kx; k1++)
        for (int k = 1; k <= kx; k++)
            for (int n = 1; n <= nx; n++)
                for (int m = 1; m <= mx; m++) ye(m, n, k) = ye(m, n, k) + xd[k][k1] * tdt(m, n, k1);
    for (int k = 1; k <= kx; k++)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) ye(m, n, k) = ye(m, n, k) + (rgas * tref[k]) * psdt(m, n);
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) yf(m, n, k) = divdt(m, n, k) + elz[Spectral::i2(m, n)] * ye(m, n, k);
    for (int k = 1; k <= kx; k++)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) divdt(m, n, k) = cplx{0.0, 0.0};
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++)
            if (m + n - 2 != 0)
                for (int k1 = 1; k1 <= kx; k1++)
                    for (int k = 1; k <= kx; k++) divdt(m, n, k) = divdt(m, n, k) + XJ(k, k1, m + n - 2) * yf(m, n, k1);
    for (int k = 1; k <= kx; k++)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) psdt(m, n) = psdt(m, n) - divdt(m, n, k) * dhsx[k];
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++)
            for (int n = 1; n <= nx; n++)
                for (int m = 1; m <= mx; m++) tdt(m, n, k) = tdt(m, n, k) + xc[k][k1] * divdt(m, n, k1);
}

// ---------------------------------------------------------------------------------------------------
// horizontal_diffusion.f90:131-152
static void hdiff3(S3 field, S3 fdt, const std::vector<double> &dmp, const std::vector<double> &dmp1) {
    for (int k = 1; k <= kx; k++)
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) {
                size_t q = Spectral::i2(m, n);
                fdt(m, n, k) = (fdt(m, n, k) - dmp[q] * field(m, n, k)) * dmp1[q];
            }
}

// time_stepping.f90:164-188 ; input/output are the two time levels of one (mx,nx) field
static void step_field_2d(const Spectral &sp, int j1, double dt, double eps, S2 lev1, S2 lev2, S2 fdt) {
    if (ix == iy * 4) sp.truncate(fdt);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            cplx o1 = lev1(m, n), o2 = lev2(m, n);
            cplx fnew = o1 + dt * fdt(m, n);
            cplx oj1 = (j1 == 1) ? o1 : o2;
            o1 = oj1 + (wil * eps) * (o1 - 2.0 * oj1 + fnew);
            if (j1 == 1) oj1 = o1;  // output(:,:,j1) aliases the freshly updated level 1
            o2 = fnew - ((1.0 - wil) * eps) * (o1 - 2.0 * oj1 + fnew);
            lev1(m, n) = o1;
            lev2(m, n) = o2;
        }
}

// time_stepping.f90:38-147
void step(State &s, int j1, int j2, double dt) {
    Spec3 vordt, divdt, tdt, trdt;
    Spec2 psdt;
    get_tendencies(s, vordt, divdt, tdt, psdt, trdt, j2);
    apply_tendencies(s, j1, dt, vordt, divdt, tdt, psdt, trdt);
}

// time_stepping.f90:78-144: horizontal diffusion, stratospheric drag and the leapfrog / RAW time integration for given
// tendencies (split off step() so that tests can feed tendencies in)
void apply_tendencies(State &s, int j1, double dt, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt) {
    const Spectral &sp = s.spec;
    Implicit &im = s.imp;
    Spec3 ctmp;

    S3 vor1 = s.s4lev(V_vor, 1), div1 = s.s4lev(V_div, 1), t1 = s.s4lev(V_t, 1), tr1 = s.s4lev(V_tr, 1);
    hdiff3(vor1, vordt, im.dmp, im.dmp1);
    hdiff3(div1, divdt, im.dmpd, im.dmp1d);
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) ctmp(m, n, k) = t1(m, n, k) + im.tcorh[Spectral::i2(m, n)] * im.tcorv[k];
    hdiff3(ctmp, tdt, im.dmp, im.dmp1);
    const double sdrag = 1.0 / (tdrs * FL(3600.0));
    for (int n = 1; n <= nx; n++) {
        vordt(1, n, 1) = vordt(1, n, 1) - sdrag * vor1(1, n, 1);
        divdt(1, n, 1) = divdt(1, n, 1) - sdrag * div1(1, n, 1);
    }
    hdiff3(vor1, vordt, im.dmps, im.dmp1s);
    hdiff3(div1, divdt, im.dmps, im.dmp1s);
    hdiff3(ctmp, tdt, im.dmps, im.dmp1s);
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) ctmp(m, n, k) = tr1(m, n, k) + im.qcorh[Spectral::i2(m, n)] * im.qcorv[k];
    hdiff3(ctmp, trdt, im.dmpd, im.dmp1d);

    double eps = (j1 == 1) ? 0.0 : rob;
    step_field_2d(sp, j1, dt, eps, s.s3lev(V_ps, 1), s.s3lev(V_ps, 2), psdt);
    const int ids[4] = {V_vor, V_div, V_t, V_tr};
    S3 fd[4] = {vordt, divdt, tdt, trdt};
    for (int q = 0; q < 4; q++)
        for (int k = 1; k <= kx; k++)
            step_field_2d(sp, j1, dt, eps, s.s4lev(ids[q], 1).slab(k), s.s4lev(ids[q], 2).slab(k), fd[q].slab(k));
}

// time_stepping.f90:13-27
void first_step(State &s) {
    s.imp.set_time_step(0.5 * delt);
    step(s, 1, 1, 0.5 * delt);
    s.imp.set_time_step(delt);
    step(s, 1, 2, delt);
    s.imp.set_time_step(2 * delt);
}

// diagnostics.f90:16-74
int check_diagnostics(State &s, int time_lev) {
    const Spectral &sp = s.spec;
    double diag[kx + 1][4];
    Spec2 temp;
    S3 vor = s.s4lev(V_vor, time_lev), div = s.s4lev(V_div, time_lev), t = s.s4lev(V_t, time_lev);
    for (int k = 1; k <= kx; k++) {
        diag[k][1] = 0.0;
        diag[k][2] = 0.0;
        diag[k][3] = F_SQRTH * t(1, 1, k).re;
        sp.laplacian_inv(vor.slab(k), temp);
        for (int m = 2; m <= mx; m++)
            for (int n = 1; n <= nx; n++) {
                cplx a = temp(m, n), b = vor(m, n, k);  // real(a * conjg(b))
                diag[k][1] = diag[k][1] - (a.re * b.re - a.im * (-b.im));
            }
        sp.laplacian_inv(div.slab(k), temp);
        for (int m = 2; m <= mx; m++)
            for (int n = 1; n <= nx; n++) {
                cplx a = temp(m, n), b = div(m, n, k);
                diag[k][2] = diag[k][2] - (a.re * b.re - a.im * (-b.im));
            }
    }
    for (int k = 1; k <= kx; k++)
        if (diag[k][1] > 500.0 || diag[k][2] > 500.0 || diag[k][3] < 180.0 || diag[k][3] > 320.0)
            return -2;  // error_codes.f90:9  E_DIAGNOSTICS_OUTSIDE_RANGE (as in the reference, NaNs pass)
    return 0;
}

}  // namespace orc
