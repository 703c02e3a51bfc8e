// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference SPEEDY hot path (aperezhortal/pySPEEDY, speedy.f90/*.f90), written
// operation-for-operation in the reference's order of evaluation, including gfortran's default-REAL(4)
// literal semantics.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
// build, link or call anything in this directory.
//
// PARITY PINNING: the reference Fortran cannot be compiled in this image (no Fortran compiler).  This oracle
// is pinned against what the reference ships: the fixture coordinates lat/lev/lon (bit-exact KAT for
// geometry.f90:89-110 + initialization.f90:85-87) and the golden fields of pyspeedy/tests/fixtures/*.nc as
// an end-to-end check (the default SST-anomaly file is missing from the mount, so that check is coarse).
// At per-transform / per-tendency granularity the reference holds no vectors: "parity unpinned" there.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/spdy_registry.h"

namespace orc {

// params.f90:18-35
constexpr int trunc_ = 30, ix = 96, iy = 24, il = 48, kx = 8, nx = 32, mx = 31, ntr = 1, t_levs = 2;
constexpr int nsteps = 36, nstrad = 3;

// A default-REAL literal widened to double, e.g. FL(0.05) == (double)0.05f  (SURVEY 7.1 rule set)
#define FL(x) ((double)(x##f))

// Constant-folded single-precision intrinsics (gfortran folds with MPFR => correctly rounded float)
constexpr double F_ATAN1 = 0x1.921fb6p-1;   // atan(1.)
constexpr double F_ASIN1 = 0x1.921fb6p+0;   // asin(1.0)
constexpr double F_SQRT2 = 0x1.6a09e6p+0;   // sqrt(2.)
constexpr double F_SQRT3 = 0x1.bb67aep+0;   // sqrt(3.)
constexpr double F_SQRTH = 0x1.6a09e6p-1;   // sqrt(0.5)
constexpr double F_LOG099 = -0x1.49544p-7;  // log(0.99)
constexpr double F_LOG1013 = 0x1.a73d3ep-7; // log(1.013)

// params.f90:33-36
const double delt = (double)(86400.0f / 36.0f);
const double rob = FL(0.05), wil = FL(0.53), alph = 0.5;

// physical_constants.f90:15-47
const double rearth = FL(6.371e+6), omega = FL(7.292e-05), grav = FL(9.81);
const double p0 = FL(1.e+5), cp = FL(1004.0), akap = (double)(2.0f / 7.0f), rgas = akap * cp;
const double alhc = FL(2501.0), alhs = FL(2801.0), sbc = FL(5.67e-8);
const double gamma_ = FL(6.0), hscale = FL(7.5), hshum = FL(2.5), refrh1 = FL(0.7);
const double thd = FL(2.4), thdd = FL(2.4), thds = FL(12.0), tdrs = (double)(24.0f * 30.0f);
// mod_radcon.f90:11-16
const double albsea = FL(0.07), albice = FL(0.60), albsn = FL(0.60), epslw = FL(0.05), emisfc = FL(0.98);

// DIAGNOSTIC SWITCHES (default off; tests/test_oracle_transcription.py only).  They replace reference *constants* by
// their exact values so that the transcription of the *algorithm* can be checked against mathematics the reference
// itself cannot satisfy: with exact_fft the FFTPACK passes must equal the DFT to rounding; with
// exact_nodes (true Gaussian latitudes instead of the REAL(4) Newton start value of geometry.f90:110) the Legendre
// pair must be an exact quadrature.  Tables built while a switch is on are not the reference's.
struct Diag {
    bool exact_fft = false, exact_nodes = false;
};
extern Diag g_diag;

struct cplx {
    double re, im;
};
inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
inline cplx operator-(cplx a) { return {-a.re, -a.im}; }
inline cplx operator*(double s, cplx a) { return {s * a.re, s * a.im}; }
inline cplx operator*(cplx a, double s) { return {a.re * s, a.im * s}; }
// (a+bi)*(0+1i) under -fcx-limited-range: (a*0 - b*1, a*1 + b*0)
inline cplx times_i(cplx a) { return {a.re * 0.0 - a.im * 1.0, a.re * 1.0 + a.im * 0.0}; }

// 1-based Fortran-order views
template <class T> struct V1 {
    T *p;
    T &operator()(int i) const { return p[i - 1]; }
};
template <class T> struct V2 {
    T *p;
    int n1;
    T &operator()(int i, int j) const { return p[(i - 1) + (size_t)n1 * (j - 1)]; }
};
template <class T> struct V3 {
    T *p;
    int n1, n2;
    T &operator()(int i, int j, int k) const { return p[(i - 1) + (size_t)n1 * ((j - 1) + (size_t)n2 * (k - 1))]; }
    V2<T> slab(int k) const { return {p + (size_t)n1 * n2 * (k - 1), n1}; }
};
template <class T> struct V4 {
    T *p;
    int n1, n2, n3;
    T &operator()(int i, int j, int k, int l) const {
        return p[(i - 1) + (size_t)n1 * ((j - 1) + (size_t)n2 * ((k - 1) + (size_t)n3 * (l - 1)))];
    }
    V2<T> slab(int k, int l) const { return {p + (size_t)n1 * n2 * ((k - 1) + (size_t)n3 * (l - 1)), n1}; }
};

typedef V2<double> G2;  // (ix, il)
typedef V3<double> G3;  // (ix, il, kx)
typedef V2<cplx> S2;    // (mx, nx)
typedef V3<cplx> S3;    // (mx, nx, kx)

struct Grid2 {
    std::vector<double> d;
    double *p;
    Grid2() : d((size_t)ix * il, 0.0), p(d.data()) {}
    Grid2(const Grid2 &) = delete;
    G2 v() { return {d.data(), ix}; }
    operator G2() { return v(); }
    double &operator()(int i, int j) { return d[(i - 1) + (size_t)ix * (j - 1)]; }
};
struct Grid3 {
    std::vector<double> d;
    int nk;
    double *p;
    explicit Grid3(int nk_ = kx) : d((size_t)ix * il * nk_, 0.0), nk(nk_), p(d.data()) {}
    Grid3(const Grid3 &) = delete;
    G3 v() { return {d.data(), ix, il}; }
    operator G3() { return v(); }
    double &operator()(int i, int j, int k) { return d[(i - 1) + (size_t)ix * ((j - 1) + (size_t)il * (k - 1))]; }
    G2 slab(int k) { return {d.data() + (size_t)ix * il * (k - 1), ix}; }
};
struct Spec2 {
    std::vector<cplx> d;
    Spec2() : d((size_t)mx * nx, cplx{0.0, 0.0}) {}
    S2 v() { return {d.data(), mx}; }
    operator S2() { return v(); }
    cplx &operator()(int m, int n) { return d[(m - 1) + (size_t)mx * (n - 1)]; }
};
struct Spec3 {
    std::vector<cplx> d;
    explicit Spec3(int nk = kx) : d((size_t)mx * nx * nk, cplx{0.0, 0.0}) {}
    S3 v() { return {d.data(), mx, nx}; }
    operator S3() { return v(); }
    cplx &operator()(int m, int n, int k) { return d[(m - 1) + (size_t)mx * ((n - 1) + (size_t)nx * (k - 1))]; }
    S2 slab(int k) { return {d.data() + (size_t)mx * nx * (k - 1), mx}; }
};

// ---------------------------------------------------------------------------------------------------
// geometry.f90:15-56
struct Geometry {
    double hsg[kx + 2], dhs[kx + 1], fsg[kx + 1], dhsr[kx + 1], fsgr[kx + 1];  // 1-based
    double radang[il + 1], coriol[il + 1], sia[il + 1], coa[il + 1], sia_half[iy + 1], coa_half[iy + 1];
    double cosg[il + 1], cosgr[il + 1], cosgr2[il + 1];
    double sigl[kx + 1], sigh[kx + 1] /* 0..kx */, grdsig[kx + 1], grdscp[kx + 1];
    double wvi[kx + 1][3];  // wvi[k][1..2]
    void initialize();
};

// legendre.f90:14-32, fourier.f90:19-29, spectral.f90:11-33
struct Spectral {
    const Geometry *geo = nullptr;
    // legendre
    std::vector<double> epsi, repsi;  // (mx+1, nx+1)
    std::vector<double> cpol;         // (2*mx, nx, iy)
    int nsh2[nx + 1];
    double wt[iy + 1];
    // fourier
    double work[ix + 1];  // wa, 1-based
    int ifac[16];         // 1-based
    // spectral
    std::vector<double> el2, elm2, el4, trfilt;                       // (mx,nx)
    std::vector<double> gradym, gradyp, uvdx, uvdym, uvdyp, vddym, vddyp;  // (mx,nx)
    double gradx[mx + 1];

    double &EPSI(int m, int n) { return epsi[(m - 1) + (size_t)(mx + 1) * (n - 1)]; }
    double &REPSI(int m, int n) { return repsi[(m - 1) + (size_t)(mx + 1) * (n - 1)]; }
    double &CPOL(int m, int n, int j) { return cpol[(m - 1) + (size_t)(2 * mx) * ((n - 1) + (size_t)nx * (j - 1))]; }
    const double &CPOL(int m, int n, int j) const {
        return cpol[(m - 1) + (size_t)(2 * mx) * ((n - 1) + (size_t)nx * (j - 1))];
    }
    static size_t i2(int m, int n) { return (m - 1) + (size_t)mx * (n - 1); }

    void initialize(const Geometry *g);
    void legendre_poly(int j, double *poly /* (mx,nx) */) const;
    // legendre.f90:130-169 / 175-221 : real-packed (2*mx, nx) <-> (2*mx, il)
    void legendre_inv(const double *in, double *out) const;
    void legendre_dir(const double *in, double *out) const;
    // fourier.f90:63-93 / 96-123
    void fourier_inv(const double *in, double *out, int kcos) const;
    void fourier_dir(const double *in, double *out) const;
    // spectral.f90:251-273
    void spec2grid(S2 vorm, G2 vorg, int kcos) const;
    void grid2spec(G2 vorg, S2 vorm) const;
    void truncate(S2 vor) const;
    void laplacian(S2 in, S2 out) const;
    void laplacian_inv(S2 in, S2 out) const;
    void vel2vort(S2 ucosm, S2 vcosm, S2 vorm, S2 divm) const;
    void vort2vel(S2 vorm, S2 divm, S2 ucosm, S2 vcosm) const;
    void grid_vel2vort(G2 ug, G2 vg, S2 vorm, S2 divm, int kcos) const;
    void gradient(S2 psi, S2 psdx, S2 psdy) const;
    void grid_filter(G2 fg1, G2 fg2) const;
};

// fftpack.f90 (N = 96 paths only)
void rffti1(int n, double *wa /*1-based*/, int *ifac /*1-based*/);
void rfftf1(int n, double *c, double *ch, const double *wa, const int *ifac);
void rfftb1(int n, double *c, double *ch, const double *wa, const int *ifac);

// horizontal_diffusion.f90:16-41 + implicit.f90:17-28
struct Implicit {
    const Geometry *geo = nullptr;
    std::vector<double> dmp, dmpd, dmps, dmp1, dmp1d, dmp1s;  // (mx,nx)
    double tcorv[kx + 1], qcorv[kx + 1];
    std::vector<cplx> tcorh, qcorh;  // (mx,nx)
    double tref[kx + 1], tref2[kx + 1], tref3[kx + 1], dhsx[kx + 1];
    double xc[kx + 1][kx + 1], xd[kx + 1][kx + 1];
    std::vector<double> xj;   // (kx,kx,mx+nx+1)
    std::vector<double> elz;  // (mx,nx)
    double &XJ(int k, int k1, int l) { return xj[(k - 1) + (size_t)kx * ((k1 - 1) + (size_t)kx * (l - 1))]; }
    void initialize(const Geometry *g);
    void set_time_step(double dt);
    void implicit_terms(S3 divdt, S3 tdt, S2 psdt);
};
void inv8(double *a /*(n,n)*/, double *y /*(n,n)*/, int n);

// model_control.f90:16-47
struct Datetime {
    int year, month, day, hour, minute;
};
struct Control {
    Datetime model_datetime, start_datetime, end_datetime;
    int imont1;
    double tmonth, tyear;
    int month_idx = 1;
    int ndaycal[13][3];
    void initialize(const Datetime &s, const Datetime &e);
    void advance_date();
    void update_forcing_params();
};

// model_state.f90 (generated in the reference): all registry variables, Fortran order, complex interleaved
struct State {
    std::vector<double> var[SPDY_NVARS];
    std::vector<float> lon, lat, lev;
    int current_step = 0;
    bool increase_co2 = false, compute_shortwave = true, land_coupling_flag = true,
         sst_anomaly_coupling_flag = true, initialized = false;
    double air_absortivity_co2 = 6.0, ablco2_ref = 0.0;
    int n_months = 0;  // sst_anom has n_months + 2 slabs
    Geometry geo;
    Spectral spec;
    Implicit imp;
    // SPPT (sppt.f90; a compile-time switch in the reference, params.f90:44: off).  The AR(1) pattern is kept per member
    // (the reference keeps ONE module-level pattern and loses it between calls: see gen_sppt in physics.cpp)
    bool sppt_on = false;
    uint64_t sppt_seed = 0, sppt_member = 0, sppt_calls = 0;  // calls of gen_sppt so far: noise counter, 0 = first AR(1) step
    std::vector<cplx> sppt_spec;   // (mx, nx, kx)
    std::vector<double> sppt_last; // (ix, il, kx) pattern of the last call (for tests)

    State();
    void alloc_sst_anom(int n_months_);
    double *p(int id) { return var[id].data(); }
    G2 g2(int id) { return {var[id].data(), ix}; }
    G3 g3(int id) { return {var[id].data(), ix, il}; }
    S2 s2(int id) { return {(cplx *)var[id].data(), mx}; }
    S3 s3(int id) { return {(cplx *)var[id].data(), mx, nx}; }
    // (mx,nx,kx,t_levs) -> level view
    S3 s4lev(int id, int tl) { return {(cplx *)var[id].data() + (size_t)mx * nx * kx * (tl - 1), mx, nx}; }
    S2 s3lev(int id, int tl) { return {(cplx *)var[id].data() + (size_t)mx * nx * (tl - 1), mx}; }
    V3<double> aux3(int id) { return {var[id].data(), ix, il}; }  // (ix,il,n)
};

// --- dynamics -------------------------------------------------------------------------------------
void initialize_geopotential(State &s);
void set_geopotential(State &s, int time_level);
void get_grid_point_tendencies(State &s, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt, int j1, int j2);
void get_spectral_tendencies(State &s, S3 divdt, S3 tdt, S2 psdt, int j2);
void get_tendencies(State &s, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt, int j2);
void step(State &s, int j1, int j2, double dt);
void apply_tendencies(State &s, int j1, double dt, S3 vordt, S3 divdt, S3 tdt, S2 psdt, S3 trdt);
void first_step(State &s);
int check_diagnostics(State &s, int time_lev);

// --- physics --------------------------------------------------------------------------------------
void get_qsat(const double *ta, const double *ps, double sig, double *qsat, int n);  // humidity.f90:44-78
void get_convection_tendencies(G2 psa, G3 se, G3 qa, G3 qsat, int *itop, G2 cbmf, G2 precnv, G3 dfse, G3 dfqa,
                               const Geometry &g);
void get_large_scale_condensation_tendencies(G2 psa, G3 qa, G3 qsat, int *itop, G2 precls, G3 dtlsc, G3 dqlsc,
                                             const Geometry &g);
void clouds(G3 qa, G3 rh, G2 precnv, G2 precls, const int *iptop, G2 gse, G2 fmask, int *icltop, G2 cloudc,
            G2 clstr, G2 qcloud_equiv);
void get_shortwave_rad_fluxes(State &s, G2 psa, G3 qa, const int *icltop, G2 cloudc, G2 clstr);
void get_zonal_average_fields(State &s, double tyear);
void radset(double *fband);
void get_downward_longwave_rad_fluxes(State &s, G3 ta, G2 fsfcd, G3 dfabs);
void get_upward_longwave_rad_fluxes(State &s, G3 ta, G2 ts, G2 fsfcd, G2 fsfcu, G2 fsfc, G2 ftop, G3 dfabs);
void get_surface_fluxes(State &s, G2 psa, G3 ua, G3 va, G3 ta, G3 qa, G3 rh, G3 phi, G2 tsea, G2 tsfc, G2 tskin,
                        G2 u0, G2 v0, G2 t0);
void set_orog_land_sfc_drag(G2 phi0, G2 forog);
void get_vertical_diffusion_tend(G3 se, G3 rh, G3 qa, G3 qsat, G3 phi, const int *icnv, G3 utenvd, G3 vtenvd,
                                 G3 ttenvd, G3 qtenvd, const Geometry &g);
void get_physical_tendencies(State &s, int j1, G3 utend, G3 vtend, G3 ttend, G3 qtend);
void gen_sppt(State &s, G3 sppt_grid);
uint64_t sppt_mix64(uint64_t z);
void physics_columns(State &s, G3 ug, G3 vg, G3 tg, G3 qg, G3 phig, G2 pslg, G3 utend, G3 vtend, G3 ttend, G3 qtend,
                     int *dbg);

// --- boundary / surface models / forcing / init -----------------------------------------------------
void initialize_boundaries(State &s);
void land_model_init(State &s);
void couple_land_atm(State &s, int day, int imont1, double tmonth);
void sea_model_init(State &s);
void couple_sea_atm(State &s, int day, const Control &c);
void set_forcing(State &s, int imode, const Datetime &dt, double tyear);
int initialize_prognostics(State &s);
int initialize_state(State &s, Control &c);
int do_single_step(State &s, Control &c);
void spectral2grid(State &s);
void grid2spectral(State &s);
void grid_filter_state(State &s);

}  // namespace orc
