// ORACLE (test infrastructure) -- geometry, FFTPACK N=96, Fourier, Legendre and spectral operators.
// Follows speedy.f90/geometry.f90, fftpack.f90, fourier.f90, legendre.f90, spectral.f90 of the reference.
#include "speedy_oracle.hpp"

namespace orc {

Diag g_diag;
// FFTPACK constants: the reference's REAL(4)-valued ones, or exact doubles under the diagnostic switch
static inline double k_tpi() { return g_diag.exact_fft ? 8.0 * atan(1.0) : 8.0 * F_ATAN1; }
static inline double k_taui() { return g_diag.exact_fft ? 0.5 * sqrt(3.0) : 0.5 * F_SQRT3; }
static inline double k_sqrt2() { return g_diag.exact_fft ? sqrt(2.0) : F_SQRT2; }
// Gaussian node i (1 = nearest the pole) by the Newton iteration of legendre.f90:224-257 (diagnostic switch only)
static double gauss_node(int i) {
    const int n = 2 * iy;
    double z = cos(3.14159265358979323846 * ((double)i - 0.25) / ((double)n + 0.5)), z1 = 2.0;
    for (int it = 0; it < 100 && fabs(z - z1) > 2.220446049250313e-16; it++) {
        double p1 = 1.0, p2 = 0.0;
        for (int j = 1; j <= n; j++) {
            double p3 = p2;
            p2 = p1;
            p1 = ((2.0 * (double)j - 1.0) * z * p2 - ((double)j - 1.0) * p3) / j;
        }
        const double pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
        z1 = z;
        z = z1 - p1 / pp;
    }
    return z;
}

// ---------------------------------------------------------------------------------------------------
// geometry.f90:61-156
void Geometry::initialize() {
    static const double hs[9] = {FL(0.000), FL(0.050), FL(0.140), FL(0.260), FL(0.420),
                                 FL(0.600), FL(0.770), FL(0.900), FL(1.000)};  // :89
    for (int k = 1; k <= kx + 1; k++) hsg[k] = hs[k - 1];
    for (int k = 1; k <= kx; k++) {  // :93-96
        dhs[k] = hsg[k + 1] - hsg[k];
        fsg[k] = 0.5 * (hsg[k + 1] + hsg[k]);
    }
    for (int k = 1; k <= kx; k++) {  // :99-102
        dhsr[k] = 0.5 / dhs[k];
        fsgr[k] = akap / (2.0 * fsg[k]);
    }
    for (int j = 1; j <= iy; j++) {  // :108-118 -- the argument and the cosine are REAL(4)
        int jj = il + 1 - j;
        float arg = 3.141592654f * ((float)j - 0.25f) / ((float)il + 0.5f);
        sia_half[j] = g_diag.exact_nodes ? gauss_node(j) : (double)cosf(arg);
        coa_half[j] = sqrt(1.0 - sia_half[j] * sia_half[j]);
        sia[j] = -sia_half[j];
        sia[jj] = sia_half[j];
        coa[j] = coa_half[j];
        coa[jj] = coa_half[j];
        radang[j] = -asin(sia_half[j]);
        radang[jj] = asin(sia_half[j]);
    }
    for (int j = 1; j <= iy; j++) {  // :121-129
        int jj = il + 1 - j;
        cosg[j] = coa_half[j];
        cosg[jj] = coa_half[j];
        cosgr[j] = 1.0 / coa_half[j];
        cosgr[jj] = 1.0 / coa_half[j];
        cosgr2[j] = 1.0 / (coa_half[j] * coa_half[j]);
        cosgr2[jj] = 1.0 / (coa_half[j] * coa_half[j]);
    }
    for (int j = 1; j <= il; j++) coriol[j] = 2.0 * omega * sia[j];  // :131
    sigh[0] = hsg[1];
    for (int k = 1; k <= kx; k++) {  // :136-141
        sigl[k] = log(fsg[k]);
        sigh[k] = hsg[k + 1];
        grdsig[k] = grav / (dhs[k] * p0);
        grdscp[k] = grdsig[k] / cp;
    }
    for (int k = 1; k <= kx - 1; k++) {  // :147-150
        wvi[k][1] = 1.0 / (sigl[k + 1] - sigl[k]);
        wvi[k][2] = (log(sigh[k]) - sigl[k]) * wvi[k][1];
    }
    wvi[kx][1] = 0.0;
    wvi[kx][2] = (F_LOG099 - sigl[kx]) * wvi[kx - 1][1];  // :153, log(0.99) is a folded REAL(4) constant
}

// ---------------------------------------------------------------------------------------------------
// fftpack.f90:1-67
void rffti1(int n, double *wa, int *ifac) {
    static const int ntryh[4] = {4, 2, 3, 5};
    int nl = n, nf = 0, j = 0, ntry = 0;
    for (;;) {
        j++;
        ntry = (j <= 4) ? ntryh[j - 1] : ntry + 2;
        for (;;) {
            int nq = nl / ntry;
            if (nl - ntry * nq != 0) break;
            nf++;
            ifac[nf + 2] = ntry;
            nl = nq;
            if (ntry == 2 && nf != 1) {
                for (int i = 2; i <= nf; i++) {
                    int ib = nf - i + 2;
                    ifac[ib + 2] = ifac[ib + 1];
                }
                ifac[3] = 2;
            }
            if (nl == 1) break;
        }
        if (nl == 1) break;
    }
    ifac[1] = n;
    ifac[2] = nf;
    const double tpi = k_tpi();  // :39  tpi = 8.*atan(1.) in REAL(4): 8*0.785398185f is exact in float
    const double argh = tpi / n;
    int is = 0, l1 = 1;
    if (nf - 1 == 0) return;
    for (int k1 = 1; k1 <= nf - 1; k1++) {
        int ip = ifac[k1 + 2], ld = 0, l2 = l1 * ip, ido = n / l2;
        for (int jj = 1; jj <= ip - 1; jj++) {
            ld += l1;
            int i = is;
            double argld = ld * argh, fi = 0.0;
            for (int ii = 3; ii <= ido; ii += 2) {
                i += 2;
                fi += 1.0;
                double arg = fi * argld;
                wa[i - 1] = cos(arg);
                wa[i] = sin(arg);
            }
            is += ido;
        }
        l1 = l2;
    }
}

// Index helpers for the FFTPACK work arrays: cc(ido,ip,l1) / ch(ido,l1,ip), all 1-based.
#define CCB(i, q, k) cc[((i)-1) + ido * (((q)-1) + ip_ * ((k)-1))]  // backward input  cc(ido,ip,l1)
#define CHB(i, k, q) ch[((i)-1) + ido * (((k)-1) + l1 * ((q)-1))]   // backward output ch(ido,l1,ip)
#define CCF(i, k, q) cc[((i)-1) + ido * (((k)-1) + l1 * ((q)-1))]   // forward input   cc(ido,l1,ip)
#define CHF(i, q, k) ch[((i)-1) + ido * (((q)-1) + ip_ * ((k)-1))]  // forward output  ch(ido,ip,l1)

// fftpack.f90:204-254
static void radb2(int ido, int l1, const double *cc, double *ch, const double *wa1 /*1-based*/) {
    const int ip_ = 2;
    for (int k = 1; k <= l1; k++) {
        CHB(1, k, 1) = CCB(1, 1, k) + CCB(ido, 2, k);
        CHB(1, k, 2) = CCB(1, 1, k) - CCB(ido, 2, k);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                CHB(i - 1, k, 1) = CCB(i - 1, 1, k) + CCB(ic - 1, 2, k);
                double tr2 = CCB(i - 1, 1, k) - CCB(ic - 1, 2, k);
                CHB(i, k, 1) = CCB(i, 1, k) - CCB(ic, 2, k);
                double ti2 = CCB(i, 1, k) + CCB(ic, 2, k);
                CHB(i - 1, k, 2) = wa1[i - 2] * tr2 - wa1[i - 1] * ti2;
                CHB(i, k, 2) = wa1[i - 2] * ti2 + wa1[i - 1] * tr2;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        CHB(ido, k, 1) = CCB(ido, 1, k) + CCB(ido, 1, k);
        CHB(ido, k, 2) = -(CCB(1, 2, k) + CCB(1, 2, k));
    }
}

// fftpack.f90:256-326
static void radb3(int ido, int l1, const double *cc, double *ch, const double *wa1, const double *wa2) {
    const int ip_ = 3;
    const double taur = -0.5, taui = k_taui();  // :268-269 REAL(4) expression .5*sqrt(3.)
    for (int k = 1; k <= l1; k++) {
        double tr2 = CCB(ido, 2, k) + CCB(ido, 2, k);
        double cr2 = CCB(1, 1, k) + taur * tr2;
        CHB(1, k, 1) = CCB(1, 1, k) + tr2;
        double ci3 = taui * (CCB(1, 3, k) + CCB(1, 3, k));
        CHB(1, k, 2) = cr2 - ci3;
        CHB(1, k, 3) = cr2 + ci3;
    }
    if (ido == 1) return;
    int idp2 = ido + 2;
    for (int k = 1; k <= l1; k++)
        for (int i = 3; i <= ido; i += 2) {
            int ic = idp2 - i;
            double tr2 = CCB(i - 1, 3, k) + CCB(ic - 1, 2, k);
            double cr2 = CCB(i - 1, 1, k) + taur * tr2;
            CHB(i - 1, k, 1) = CCB(i - 1, 1, k) + tr2;
            double ti2 = CCB(i, 3, k) - CCB(ic, 2, k);
            double ci2 = CCB(i, 1, k) + taur * ti2;
            CHB(i, k, 1) = CCB(i, 1, k) + ti2;
            double cr3 = taui * (CCB(i - 1, 3, k) - CCB(ic - 1, 2, k));
            double ci3 = taui * (CCB(i, 3, k) + CCB(ic, 2, k));
            double dr2 = cr2 - ci3, dr3 = cr2 + ci3, di2 = ci2 + cr3, di3 = ci2 - cr3;
            CHB(i - 1, k, 2) = wa1[i - 2] * dr2 - wa1[i - 1] * di2;
            CHB(i, k, 2) = wa1[i - 2] * di2 + wa1[i - 1] * dr2;
            CHB(i - 1, k, 3) = wa2[i - 2] * dr3 - wa2[i - 1] * di3;
            CHB(i, k, 3) = wa2[i - 2] * di3 + wa2[i - 1] * dr3;
        }
}

// fftpack.f90:328-424
static void radb4(int ido, int l1, const double *cc, double *ch, const double *wa1, const double *wa2,
                  const double *wa3) {
    const int ip_ = 4;
    const double sqrt2 = k_sqrt2();  // :341 sqrt(2.) REAL(4)
    for (int k = 1; k <= l1; k++) {
        double tr1 = CCB(1, 1, k) - CCB(ido, 4, k);
        double tr2 = CCB(1, 1, k) + CCB(ido, 4, k);
        double tr3 = CCB(ido, 2, k) + CCB(ido, 2, k);
        double tr4 = CCB(1, 3, k) + CCB(1, 3, k);
        CHB(1, k, 1) = tr2 + tr3;
        CHB(1, k, 2) = tr1 - tr4;
        CHB(1, k, 3) = tr2 - tr3;
        CHB(1, k, 4) = tr1 + tr4;
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double ti1 = CCB(i, 1, k) + CCB(ic, 4, k);
                double ti2 = CCB(i, 1, k) - CCB(ic, 4, k);
                double ti3 = CCB(i, 3, k) - CCB(ic, 2, k);
                double tr4 = CCB(i, 3, k) + CCB(ic, 2, k);
                double tr1 = CCB(i - 1, 1, k) - CCB(ic - 1, 4, k);
                double tr2 = CCB(i - 1, 1, k) + CCB(ic - 1, 4, k);
                double ti4 = CCB(i - 1, 3, k) - CCB(ic - 1, 2, k);
                double tr3 = CCB(i - 1, 3, k) + CCB(ic - 1, 2, k);
                CHB(i - 1, k, 1) = tr2 + tr3;
                double cr3 = tr2 - tr3;
                CHB(i, k, 1) = ti2 + ti3;
                double ci3 = ti2 - ti3;
                double cr2 = tr1 - tr4, cr4 = tr1 + tr4, ci2 = ti1 + ti4, ci4 = ti1 - ti4;
                CHB(i - 1, k, 2) = wa1[i - 2] * cr2 - wa1[i - 1] * ci2;
                CHB(i, k, 2) = wa1[i - 2] * ci2 + wa1[i - 1] * cr2;
                CHB(i - 1, k, 3) = wa2[i - 2] * cr3 - wa2[i - 1] * ci3;
                CHB(i, k, 3) = wa2[i - 2] * ci3 + wa2[i - 1] * cr3;
                CHB(i - 1, k, 4) = wa3[i - 2] * cr4 - wa3[i - 1] * ci4;
                CHB(i, k, 4) = wa3[i - 2] * ci4 + wa3[i - 1] * cr4;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        double ti1 = CCB(1, 2, k) + CCB(1, 4, k);
        double ti2 = CCB(1, 4, k) - CCB(1, 2, k);
        double tr1 = CCB(ido, 1, k) - CCB(ido, 3, k);
        double tr2 = CCB(ido, 1, k) + CCB(ido, 3, k);
        CHB(ido, k, 1) = tr2 + tr2;
        CHB(ido, k, 2) = sqrt2 * (tr1 - ti1);
        CHB(ido, k, 3) = ti2 + ti2;
        CHB(ido, k, 4) = -sqrt2 * (tr1 + ti1);
    }
}

// fftpack.f90:722-772
static void radf2(int ido, int l1, const double *cc, double *ch, const double *wa1) {
    const int ip_ = 2;
    for (int k = 1; k <= l1; k++) {
        CHF(1, 1, k) = CCF(1, k, 1) + CCF(1, k, 2);
        CHF(ido, 2, k) = CCF(1, k, 1) - CCF(1, k, 2);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double tr2 = wa1[i - 2] * CCF(i - 1, k, 2) + wa1[i - 1] * CCF(i, k, 2);
                double ti2 = wa1[i - 2] * CCF(i, k, 2) - wa1[i - 1] * CCF(i - 1, k, 2);
                CHF(i, 1, k) = CCF(i, k, 1) + ti2;
                CHF(ic, 2, k) = ti2 - CCF(i, k, 1);
                CHF(i - 1, 1, k) = CCF(i - 1, k, 1) + tr2;
                CHF(ic - 1, 2, k) = CCF(i - 1, k, 1) - tr2;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        CHF(1, 2, k) = -CCF(ido, k, 2);
        CHF(ido, 1, k) = CCF(ido, k, 1);
    }
}

// fftpack.f90:774-842
static void radf3(int ido, int l1, const double *cc, double *ch, const double *wa1, const double *wa2) {
    const int ip_ = 3;
    const double taur = -0.5, taui = k_taui();
    for (int k = 1; k <= l1; k++) {
        double cr2 = CCF(1, k, 2) + CCF(1, k, 3);
        CHF(1, 1, k) = CCF(1, k, 1) + cr2;
        CHF(1, 3, k) = taui * (CCF(1, k, 3) - CCF(1, k, 2));
        CHF(ido, 2, k) = CCF(1, k, 1) + taur * cr2;
    }
    if (ido == 1) return;
    int idp2 = ido + 2;
    for (int k = 1; k <= l1; k++)
        for (int i = 3; i <= ido; i += 2) {
            int ic = idp2 - i;
            double dr2 = wa1[i - 2] * CCF(i - 1, k, 2) + wa1[i - 1] * CCF(i, k, 2);
            double di2 = wa1[i - 2] * CCF(i, k, 2) - wa1[i - 1] * CCF(i - 1, k, 2);
            double dr3 = wa2[i - 2] * CCF(i - 1, k, 3) + wa2[i - 1] * CCF(i, k, 3);
            double di3 = wa2[i - 2] * CCF(i, k, 3) - wa2[i - 1] * CCF(i - 1, k, 3);
            double cr2 = dr2 + dr3, ci2 = di2 + di3;
            CHF(i - 1, 1, k) = CCF(i - 1, k, 1) + cr2;
            CHF(i, 1, k) = CCF(i, k, 1) + ci2;
            double tr2 = CCF(i - 1, k, 1) + taur * cr2;
            double ti2 = CCF(i, k, 1) + taur * ci2;
            double tr3 = taui * (di2 - di3);
            double ti3 = taui * (dr3 - dr2);
            CHF(i - 1, 3, k) = tr2 + tr3;
            CHF(ic - 1, 2, k) = tr2 - tr3;
            CHF(i, 3, k) = ti2 + ti3;
            CHF(ic, 2, k) = ti3 - ti2;
        }
}

// fftpack.f90:844-936
static void radf4(int ido, int l1, const double *cc, double *ch, const double *wa1, const double *wa2,
                  const double *wa3) {
    const int ip_ = 4;
    const double hsqt2 = 0.5 * k_sqrt2();  // :857 .5*sqrt(2.) REAL(4) (exact halving)
    for (int k = 1; k <= l1; k++) {
        double tr1 = CCF(1, k, 2) + CCF(1, k, 4);
        double tr2 = CCF(1, k, 1) + CCF(1, k, 3);
        CHF(1, 1, k) = tr1 + tr2;
        CHF(ido, 4, k) = tr2 - tr1;
        CHF(ido, 2, k) = CCF(1, k, 1) - CCF(1, k, 3);
        CHF(1, 3, k) = CCF(1, k, 4) - CCF(1, k, 2);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double cr2 = wa1[i - 2] * CCF(i - 1, k, 2) + wa1[i - 1] * CCF(i, k, 2);
                double ci2 = wa1[i - 2] * CCF(i, k, 2) - wa1[i - 1] * CCF(i - 1, k, 2);
                double cr3 = wa2[i - 2] * CCF(i - 1, k, 3) + wa2[i - 1] * CCF(i, k, 3);
                double ci3 = wa2[i - 2] * CCF(i, k, 3) - wa2[i - 1] * CCF(i - 1, k, 3);
                double cr4 = wa3[i - 2] * CCF(i - 1, k, 4) + wa3[i - 1] * CCF(i, k, 4);
                double ci4 = wa3[i - 2] * CCF(i, k, 4) - wa3[i - 1] * CCF(i - 1, k, 4);
                double tr1 = cr2 + cr4, tr4 = cr4 - cr2, ti1 = ci2 + ci4, ti4 = ci2 - ci4;
                double ti2 = CCF(i, k, 1) + ci3, ti3 = CCF(i, k, 1) - ci3;
                double tr2 = CCF(i - 1, k, 1) + cr3, tr3 = CCF(i - 1, k, 1) - cr3;
                CHF(i - 1, 1, k) = tr1 + tr2;
                CHF(ic - 1, 4, k) = tr2 - tr1;
                CHF(i, 1, k) = ti1 + ti2;
                CHF(ic, 4, k) = ti1 - ti2;
                CHF(i - 1, 3, k) = ti4 + tr3;
                CHF(ic - 1, 2, k) = tr3 - ti4;
                CHF(i, 3, k) = tr4 + ti3;
                CHF(ic, 2, k) = tr4 - ti3;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        double ti1 = -hsqt2 * (CCF(ido, k, 2) + CCF(ido, k, 4));
        double tr1 = hsqt2 * (CCF(ido, k, 2) - CCF(ido, k, 4));
        CHF(ido, 1, k) = tr1 + CCF(ido, k, 1);
        CHF(ido, 3, k) = CCF(ido, k, 1) - tr1;
        CHF(1, 2, k) = ti1 - CCF(ido, k, 3);
        CHF(1, 4, k) = ti1 + CCF(ido, k, 3);
    }
}

// fftpack.f90:69-134  (c, ch 0-based buffers of n; wa, ifac 1-based)
void rfftb1(int n, double *c, double *ch, const double *wa, const int *ifac) {
    int nf = ifac[2], na = 0, l1 = 1, iw = 1;
    for (int k1 = 1; k1 <= nf; k1++) {
        int ip = ifac[k1 + 2], l2 = ip * l1, ido = n / l2;
        const double *in = na ? ch : c;
        double *out = na ? c : ch;
        // wa(iw) as a 1-based array => pointer wa + iw - 1 with [1] == wa(iw)
        const double *w1 = wa + iw - 1, *w2 = w1 + ido, *w3 = w2 + ido;
        if (ip == 4)
            radb4(ido, l1, in, out, w1, w2, w3);
        else if (ip == 2)
            radb2(ido, l1, in, out, w1);
        else
            radb3(ido, l1, in, out, w1, w2);
        na = 1 - na;
        l1 = l2;
        iw += (ip - 1) * ido;
    }
    if (na == 0) return;
    for (int i = 0; i < n; i++) c[i] = ch[i];
}

// fftpack.f90:136-202
void rfftf1(int n, double *c, double *ch, const double *wa, const int *ifac) {
    int nf = ifac[2], na = 1, l2 = n, iw = n;
    for (int k1 = 1; k1 <= nf; k1++) {
        int kh = nf - k1, ip = ifac[kh + 3], l1 = l2 / ip, ido = n / l2;
        iw -= (ip - 1) * ido;
        na = 1 - na;
        const double *in = na ? ch : c;
        double *out = na ? c : ch;
        const double *w1 = wa + iw - 1, *w2 = w1 + ido, *w3 = w2 + ido;
        if (ip == 4)
            radf4(ido, l1, in, out, w1, w2, w3);
        else if (ip == 2)
            radf2(ido, l1, in, out, w1);
        else
            radf3(ido, l1, in, out, w1, w2);
        l2 = l1;
    }
    if (na == 1) return;
    for (int i = 0; i < n; i++) c[i] = ch[i];
}

// ---------------------------------------------------------------------------------------------------
// legendre.f90:224-257
static void get_weights(double *w /*1-based iy*/) {
    const int n = 2 * iy;
    double z1 = 2.0, pp = 0.0;
    for (int i = 1; i <= iy; i++) {
        double z = cos(3.141592654 * ((double)i - 0.25) / ((double)n + 0.5));
        while (fabs(z - z1) > 2.220446049250313e-16) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * (double)j - 1.0) * z * p2 - ((double)j - 1.0) * p3) / j;
            }
            pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
            z1 = z;
            z = z1 - p1 / pp;
        }
        w[i] = 2.0 / ((1.0 - z * z) * (pp * pp));
    }
}

// legendre.f90:260-307
void Spectral::legendre_poly(int j, double *poly) const {
    const double small = FL(1.e-30);
    std::vector<double> alp((size_t)(mx + 1) * nx);
    double consq[mx + 1];
    auto ALP = [&](int m, int n) -> double & { return alp[(m - 1) + (size_t)(mx + 1) * (n - 1)]; };
    auto EP = [&](int m, int n) { return epsi[(m - 1) + (size_t)(mx + 1) * (n - 1)]; };
    auto REP = [&](int m, int n) { return repsi[(m - 1) + (size_t)(mx + 1) * (n - 1)]; };
    double y = geo->coa_half[j], x = geo->sia_half[j];
    for (int m = 1; m <= mx; m++)  // :277 all REAL(4), sqrtf
        consq[m] = g_diag.exact_nodes ? sqrt(0.5 * (2.0 * m + 1.0) / m)
                                      : (double)sqrtf(0.5f * (2.0f * (float)m + 1.0f) / (float)m);
    ALP(1, 1) = g_diag.exact_nodes ? sqrt(0.5) : F_SQRTH;  // :281 sqrt(0.5) folded REAL(4)
    for (int m = 2; m <= mx + 1; m++) ALP(m, 1) = consq[m - 1] * y * ALP(m - 1, 1);
    for (int m = 1; m <= mx + 1; m++) ALP(m, 2) = (x * ALP(m, 1)) * REP(m, 2);
    for (int n = 3; n <= nx; n++)
        for (int m = 1; m <= mx + 1; m++)
            ALP(m, n) = (x * ALP(m, n - 1) - EP(m, n - 1) * ALP(m, n - 2)) * REP(m, n);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx + 1; m++)
            if (fabs(ALP(m, n)) <= small) ALP(m, n) = 0.0;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) poly[(m - 1) + (size_t)mx * (n - 1)] = ALP(m, n);
}

// legendre.f90:38-112, fourier.f90:38-53, spectral.f90:39-116
void Spectral::initialize(const Geometry *g) {
    geo = g;
    cpol.assign((size_t)2 * mx * nx * iy, 0.0);
    epsi.assign((size_t)(mx + 1) * (nx + 1), 0.0);
    repsi.assign((size_t)(mx + 1) * (nx + 1), 0.0);
    get_weights(wt);
    for (int n = 1; n <= nx; n++) {  // :68-77
        nsh2[n] = 0;
        for (int m = 1; m <= mx; m++) {
            int wavenum_tot = (m - 1) + n - 1;
            if (wavenum_tot <= trunc_ + 1 || ix != 4 * iy) nsh2[n] += 2;
        }
    }
    for (int m = 1; m <= mx + 1; m++)  // :79-95
        for (int n = 1; n <= nx + 1; n++) {
            float fm = (float)(m - 1), fl = (float)(n + m - 2);
            double emm2 = (double)(fm * fm), ell2 = (double)(fl * fl);
            if (n == nx + 1)
                EPSI(m, n) = 0.0;
            else if (n == 1 && m == 1)
                EPSI(m, n) = 0.0;
            else
                EPSI(m, n) = sqrt((ell2 - emm2) / (4.0 * ell2 - 1.0));
            REPSI(m, n) = 0.0;
            if (EPSI(m, n) > 0.0) REPSI(m, n) = 1.0 / EPSI(m, n);
        }
    std::vector<double> poly((size_t)mx * nx);
    for (int j = 1; j <= iy; j++) {  // :98-108
        legendre_poly(j, poly.data());
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) {
                CPOL(2 * m - 1, n, j) = poly[(m - 1) + (size_t)mx * (n - 1)];
                CPOL(2 * m, n, j) = poly[(m - 1) + (size_t)mx * (n - 1)];
            }
    }
    // fourier.f90:49-52
    for (int i = 0; i <= ix; i++) work[i] = 0.0;
    for (int i = 0; i < 16; i++) ifac[i] = 0;
    rffti1(ix, work, ifac);

    // spectral.f90:39-116
    size_t ns = (size_t)mx * nx;
    el2.assign(ns, 0.0), elm2.assign(ns, 0.0), el4.assign(ns, 0.0), trfilt.assign(ns, 0.0);
    gradym.assign(ns, 0.0), gradyp.assign(ns, 0.0), uvdx.assign(ns, 0.0), uvdym.assign(ns, 0.0);
    uvdyp.assign(ns, 0.0), vddym.assign(ns, 0.0), vddyp.assign(ns, 0.0);
    const double rearth2 = rearth * rearth;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            int wt_ = (m - 1) + n - 1;
            el2[i2(m, n)] = (double)(float)(wt_ * (wt_ + 1)) / rearth2;
            el4[i2(m, n)] = el2[i2(m, n)] * el2[i2(m, n)];
            trfilt[i2(m, n)] = (wt_ <= trunc_) ? 1.0 : 0.0;
        }
    elm2[i2(1, 1)] = 0.0;
    for (int n = 1; n <= nx; n++)
        for (int m = 2; m <= mx; m++) elm2[i2(m, n)] = 1.0 / el2[i2(m, n)];
    for (int n = 2; n <= nx; n++) elm2[i2(1, n)] = 1.0 / el2[i2(1, n)];
    for (int m = 1; m <= mx; m++)
        for (int n = 1; n <= nx; n++) {
            int m1 = m - 1, m2 = m1 + 1;
            double el1 = (double)(float)((m - 1) + n - 1);
            if (n == 1) {
                gradx[m] = (double)(float)m1 / rearth;
                uvdx[i2(m, 1)] = -rearth / (double)(float)(m1 + 1);
                uvdym[i2(m, 1)] = 0.0;
                vddym[i2(m, 1)] = 0.0;
            } else {
                uvdx[i2(m, n)] = -rearth * (double)(float)m1 / (el1 * (el1 + 1));
                gradym[i2(m, n)] = (el1 - 1.0) * EPSI(m2, n) / rearth;
                uvdym[i2(m, n)] = -rearth * EPSI(m2, n) / el1;
                vddym[i2(m, n)] = (el1 + 1) * EPSI(m2, n) / rearth;
            }
            gradyp[i2(m, n)] = (el1 + 2.0) * EPSI(m2, n + 1) / rearth;
            uvdyp[i2(m, n)] = -rearth * EPSI(m2, n + 1) / (el1 + 1.0);
            vddyp[i2(m, n)] = el1 * EPSI(m2, n + 1) / rearth;
        }
}

// legendre.f90:130-169
void Spectral::legendre_inv(const double *in, double *out) const {
    const int m2 = 2 * mx;
    double even[2 * mx + 1], odd[2 * mx + 1];
    for (int j = 1; j <= iy; j++) {
        int j1 = il + 1 - j;
        for (int m = 1; m <= m2; m++) even[m] = 0.0, odd[m] = 0.0;
        for (int n = 1; n <= nx; n += 2)
            for (int m = 1; m <= nsh2[n]; m++) even[m] = even[m] + in[(m - 1) + m2 * (n - 1)] * CPOL(m, n, j);
        for (int n = 2; n <= nx; n += 2)
            for (int m = 1; m <= nsh2[n]; m++) odd[m] = odd[m] + in[(m - 1) + m2 * (n - 1)] * CPOL(m, n, j);
        for (int m = 1; m <= m2; m++) {
            out[(m - 1) + m2 * (j1 - 1)] = even[m] + odd[m];
            out[(m - 1) + m2 * (j - 1)] = even[m] - odd[m];
        }
    }
}

// legendre.f90:175-221
void Spectral::legendre_dir(const double *in, double *out) const {
    const int m2 = 2 * mx;
    std::vector<double> even((size_t)m2 * iy), odd((size_t)m2 * iy);
    for (int i = 0; i < m2 * nx; i++) out[i] = 0.0;
    for (int j = 1; j <= iy; j++) {
        int j1 = il + 1 - j;
        for (int m = 1; m <= m2; m++) {
            even[(m - 1) + m2 * (j - 1)] = (in[(m - 1) + m2 * (j1 - 1)] + in[(m - 1) + m2 * (j - 1)]) * wt[j];
            odd[(m - 1) + m2 * (j - 1)] = (in[(m - 1) + m2 * (j1 - 1)] - in[(m - 1) + m2 * (j - 1)]) * wt[j];
        }
    }
    for (int n = 1; n <= trunc_ + 1; n += 2)
        for (int m = 1; m <= nsh2[n]; m++) {
            double s = 0.0;  // dot_product: sequential sum j = 1..iy
            for (int j = 1; j <= iy; j++) s += CPOL(m, n, j) * even[(m - 1) + m2 * (j - 1)];
            out[(m - 1) + m2 * (n - 1)] = s;
        }
    for (int n = 2; n <= trunc_ + 1; n += 2)
        for (int m = 1; m <= nsh2[n]; m++) {
            double s = 0.0;
            for (int j = 1; j <= iy; j++) s += CPOL(m, n, j) * odd[(m - 1) + m2 * (j - 1)];
            out[(m - 1) + m2 * (n - 1)] = s;
        }
}

// fourier.f90:63-93
void Spectral::fourier_inv(const double *in, double *out, int kcos) const {
    const int m2 = 2 * mx;
    double fvar[ix], ch[ix];
    for (int j = 1; j <= il; j++) {
        fvar[0] = in[0 + m2 * (j - 1)];
        for (int m = 3; m <= m2; m++) fvar[m - 2] = in[(m - 1) + m2 * (j - 1)];
        for (int m = m2; m <= ix; m++) fvar[m - 1] = 0.0;
        rfftb1(ix, fvar, ch, work, ifac);
        if (kcos == 1)
            for (int i = 0; i < ix; i++) out[i + ix * (j - 1)] = fvar[i];
        else
            for (int i = 0; i < ix; i++) out[i + ix * (j - 1)] = fvar[i] * geo->cosgr[j];
    }
}

// fourier.f90:96-123
void Spectral::fourier_dir(const double *in, double *out) const {
    const int m2 = 2 * mx;
    double fvar[ix], ch[ix];
    const double scale = (double)(1.0f / (float)ix);  // :113 REAL(4) division
    for (int j = 1; j <= il; j++) {
        for (int i = 0; i < ix; i++) fvar[i] = in[i + ix * (j - 1)];
        rfftf1(ix, fvar, ch, work, ifac);
        out[0 + m2 * (j - 1)] = fvar[0] * scale;
        out[1 + m2 * (j - 1)] = 0.0;
        for (int m = 3; m <= m2; m++) out[(m - 1) + m2 * (j - 1)] = fvar[m - 2] * scale;
    }
}

// spectral.f90:251-273
void Spectral::spec2grid(S2 vorm, G2 vorg, int kcos) const {
    double four[2 * mx * il];
    legendre_inv((const double *)vorm.p, four);
    fourier_inv(four, vorg.p, kcos);
}
void Spectral::grid2spec(G2 vorg, S2 vorm) const {
    double four[2 * mx * il];
    fourier_dir(vorg.p, four);
    legendre_dir(four, (double *)vorm.p);
}

// spectral.f90:134-155
void Spectral::truncate(S2 vor) const {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) vor(m, n) = vor(m, n) * trfilt[i2(m, n)];
}
void Spectral::laplacian(S2 in, S2 out) const {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) out(m, n) = (-in(m, n)) * el2[i2(m, n)];
}
void Spectral::laplacian_inv(S2 in, S2 out) const {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) out(m, n) = (-in(m, n)) * elm2[i2(m, n)];
}

// spectral.f90:160-186
void Spectral::vel2vort(S2 ucosm, S2 vcosm, S2 vorm, S2 divm) const {
    Spec2 zc, zp;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            zp(m, n) = times_i(gradx[m] * ucosm(m, n));
            zc(m, n) = times_i(gradx[m] * vcosm(m, n));
        }
    for (int m = 1; m <= mx; m++) {
        vorm(m, 1) = zc(m, 1) - vddyp[i2(m, 1)] * ucosm(m, 2);
        vorm(m, nx) = vddym[i2(m, nx)] * ucosm(m, trunc_ + 1);
        divm(m, 1) = zp(m, 1) + vddyp[i2(m, 1)] * vcosm(m, 2);
        divm(m, nx) = (-vddym[i2(m, nx)]) * vcosm(m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++) {
            vorm(m, n) = vddym[i2(m, n)] * ucosm(m, n - 1) - vddyp[i2(m, n)] * ucosm(m, n + 1) + zc(m, n);
            divm(m, n) = (-vddym[i2(m, n)]) * vcosm(m, n - 1) + vddyp[i2(m, n)] * vcosm(m, n + 1) + zp(m, n);
        }
}

// spectral.f90:190-214
void Spectral::vort2vel(S2 vorm, S2 divm, S2 ucosm, S2 vcosm) const {
    Spec2 zc, zp;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            zp(m, n) = times_i(uvdx[i2(m, n)] * vorm(m, n));
            zc(m, n) = times_i(uvdx[i2(m, n)] * divm(m, n));
        }
    for (int m = 1; m <= mx; m++) {
        ucosm(m, 1) = zc(m, 1) - uvdyp[i2(m, 1)] * vorm(m, 2);
        ucosm(m, nx) = uvdym[i2(m, nx)] * vorm(m, trunc_ + 1);
        vcosm(m, 1) = zp(m, 1) + uvdyp[i2(m, 1)] * divm(m, 2);
        vcosm(m, nx) = (-uvdym[i2(m, nx)]) * divm(m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++) {
            vcosm(m, n) = (-uvdym[i2(m, n)]) * divm(m, n - 1) + uvdyp[i2(m, n)] * divm(m, n + 1) + zp(m, n);
            ucosm(m, n) = uvdym[i2(m, n)] * vorm(m, n - 1) - uvdyp[i2(m, n)] * vorm(m, n + 1) + zc(m, n);
        }
}

// spectral.f90:218-248
void Spectral::grid_vel2vort(G2 ug, G2 vg, S2 vorm, S2 divm, int kcos) const {
    Grid2 ug1, vg1;
    Spec2 specu, specv;
    for (int j = 1; j <= il; j++)
        for (int i = 1; i <= ix; i++) {
            double c = (kcos == 2) ? geo->cosgr[j] : geo->cosgr2[j];
            ug1(i, j) = ug(i, j) * c;
            vg1(i, j) = vg(i, j) * c;
        }
    grid2spec(ug1, specu);
    grid2spec(vg1, specv);
    vel2vort(specu, specv, vorm, divm);
}

// spectral.f90:275-296
void Spectral::gradient(S2 psi, S2 psdx, S2 psdy) const {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) psdx(m, n) = times_i(gradx[m] * psi(m, n));
    for (int m = 1; m <= mx; m++) {
        psdy(m, 1) = gradyp[i2(m, 1)] * psi(m, 2);
        psdy(m, nx) = (-gradym[i2(m, nx)]) * psi(m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++)
            psdy(m, n) = (-gradym[i2(m, n)]) * psi(m, n - 1) + gradyp[i2(m, n)] * psi(m, n + 1);
}

// spectral.f90:299-317
void Spectral::grid_filter(G2 fg1, G2 fg2) const {
    Spec2 fsp;
    grid2spec(fg1, fsp);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++)
            if (m + n - 2 > trunc_) fsp(m, n) = cplx{0.0, 0.0};
    spec2grid(fsp, fg2, 1);
}

}  // namespace orc
