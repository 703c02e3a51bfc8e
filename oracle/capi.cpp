// ORACLE (test infrastructure) -- C entry points for ctypes (tests/, smoke() and bench.py CPU legs only).
#include <omp.h>

#include <cstdio>
#include <string>

#include "speedy_oracle.hpp"

using namespace orc;

namespace {
State *g_tables = nullptr;
State &tables() {  // module tables only (geometry, spectral, implicit with the leapfrog time step 2*delt)
    State *&t = g_tables;
    if (!t) {
        t = new State();
        t->geo.initialize();
        t->spec.initialize(&t->geo);
        t->imp.initialize(&t->geo);
        t->imp.set_time_step(2 * delt);
        initialize_geopotential(*t);
        radset(t->p(V_fband));
    }
    return *t;
}
size_t var_count(const State &s, int v) { return s.var[v].size(); }
}  // namespace

extern "C" {

// diagnostic switches of speedy_oracle.hpp (Diag); the shared tables are rebuilt on the next use
void orc_set_diag(int exact_fft, int exact_nodes) {
    g_diag.exact_fft = exact_fft != 0, g_diag.exact_nodes = exact_nodes != 0;
    delete g_tables;
    g_tables = nullptr;
}

void *orc_state_create() { return new State(); }
void *orc_state_clone(void *p) {
    State *s = new State(*(State *)p);
    s->spec.geo = &s->geo;  // re-point the module back-references of the copy
    s->imp.geo = &s->geo;
    return s;
}
void *orc_control_clone(void *p) { return new Control(*(Control *)p); }
void orc_state_destroy(void *p) { delete (State *)p; }
void orc_alloc_sst_anom(void *p, int n_months) { ((State *)p)->alloc_sst_anom(n_months); }

void *orc_control_create(const int *s, const int *e) {
    Control *c = new Control();
    c->initialize(Datetime{s[0], s[1], s[2], s[3], s[4]}, Datetime{e[0], e[1], e[2], e[3], e[4]});
    return c;
}
void orc_control_destroy(void *p) { delete (Control *)p; }
void orc_control_date(void *p, int *out) {
    const Datetime &d = ((Control *)p)->model_datetime;
    out[0] = d.year, out[1] = d.month, out[2] = d.day, out[3] = d.hour, out[4] = d.minute;
}
void orc_control_forcing(void *p, double *out) {
    Control *c = (Control *)p;
    out[0] = c->tmonth, out[1] = c->tyear, out[2] = c->imont1, out[3] = c->month_idx;
}

int orc_init(void *st, void *ctl) { return initialize_state(*(State *)st, *(Control *)ctl); }
int orc_step(void *st, void *ctl) { return do_single_step(*(State *)st, *(Control *)ctl); }
// registry/templates/speedy_driver.f90.j2:58-79 : !$OMP PARALLEL DO SCHEDULE(dynamic) over members
void orc_parallel_step(void **st, void **ctl, int *err, int n, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int m = 0; m < n; m++) err[m] = do_single_step(*(State *)st[m], *(Control *)ctl[m]);
}
int orc_max_threads() { return omp_get_max_threads(); }
int orc_check(void *st) { return check_diagnostics(*(State *)st, 1); }
void orc_spectral2grid(void *st) { spectral2grid(*(State *)st); }
void orc_grid2spectral(void *st) { grid2spectral(*(State *)st); }
void orc_grid_filter(void *st) { grid_filter_state(*(State *)st); }

// raw accessors: f8/c16 -> doubles, f4 -> floats, i4/b1 -> int32
int orc_shape(void *st, int v, int *dims, int *ndim) {
    State &s = *(State *)st;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    *ndim = d.ndim;
    for (int q = 0; q < d.ndim; q++) dims[q] = d.dims[q] < 0 ? (s.var[v].empty() ? 0 : s.n_months + 2) : d.dims[q];
    return 0;
}
static int *scalar_i(State &s, int v, int &tmp, bool set, int val) {
    (void)tmp;
    switch (v) {
        case V_current_step: if (set) s.current_step = val; tmp = s.current_step; break;
        case V_increase_co2: if (set) s.increase_co2 = val != 0; tmp = s.increase_co2; break;
        case V_compute_shortwave: if (set) s.compute_shortwave = val != 0; tmp = s.compute_shortwave; break;
        case V_land_coupling_flag: if (set) s.land_coupling_flag = val != 0; tmp = s.land_coupling_flag; break;
        case V_sst_anomaly_coupling_flag:
            if (set) s.sst_anomaly_coupling_flag = val != 0;
            tmp = s.sst_anomaly_coupling_flag;
            break;
        default: return nullptr;
    }
    return &tmp;
}
int orc_get(void *st, int v, void *dst, size_t bytes) {
    State &s = *(State *)st;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    if (d.ndim == 0) {
        int tmp;
        if (d.kind == SPDY_F8) {
            double x = (v == V_air_absortivity_co2) ? s.air_absortivity_co2 : s.ablco2_ref;
            if (bytes != 8) return -1;
            memcpy(dst, &x, 8);
            return 0;
        }
        if (!scalar_i(s, v, tmp, false, 0) || bytes != 4) return -1;
        memcpy(dst, &tmp, 4);
        return 0;
    }
    if (d.kind == SPDY_F4) {
        std::vector<float> &f = (v == V_lon) ? s.lon : (v == V_lat) ? s.lat : s.lev;
        if (bytes != f.size() * 4) return -1;
        memcpy(dst, f.data(), bytes);
        return 0;
    }
    if (bytes != var_count(s, v) * 8) return -1;
    memcpy(dst, s.var[v].data(), bytes);
    return 0;
}
int orc_set(void *st, int v, const void *src, size_t bytes) {
    State &s = *(State *)st;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    if (d.ndim == 0) {
        int tmp;
        if (d.kind == SPDY_F8) {
            if (bytes != 8) return -1;
            double x;
            memcpy(&x, src, 8);
            if (v == V_air_absortivity_co2) s.air_absortivity_co2 = x; else s.ablco2_ref = x;
            return 0;
        }
        if (bytes != 4) return -1;
        int val;
        memcpy(&val, src, 4);
        return scalar_i(s, v, tmp, true, val) ? 0 : -1;
    }
    if (d.kind == SPDY_F4) {
        std::vector<float> &f = (v == V_lon) ? s.lon : (v == V_lat) ? s.lat : s.lev;
        if (bytes != f.size() * 4) return -1;
        memcpy(f.data(), src, bytes);
        return 0;
    }
    if (bytes != var_count(s, v) * 8) return -1;
    memcpy(s.var[v].data(), src, bytes);
    return 0;
}
// module tables of one state (tcorh/qcorh are per-member, set by the daily forcing)
int orc_get_corh(void *st, double *tcorh, double *qcorh) {
    State &s = *(State *)st;
    memcpy(tcorh, s.imp.tcorh.data(), sizeof(cplx) * mx * nx);
    memcpy(qcorh, s.imp.qcorh.data(), sizeof(cplx) * mx * nx);
    return 0;
}

// ---- shared tables by name (bit-exact comparison against the product's host table generator) ----------
int orc_table(const char *name, double *dst, int cap) {
    State &t = tables();
    const Geometry &g = t.geo;
    Spectral &sp = t.spec;
    Implicit &im = t.imp;
    std::string n(name);
    std::vector<double> out;
    auto put1 = [&](const double *a, int lo, int hi) { for (int i = lo; i <= hi; i++) out.push_back(a[i]); };
    if (n == "hsg") put1(g.hsg, 1, kx + 1);
    else if (n == "dhs") put1(g.dhs, 1, kx);
    else if (n == "fsg") put1(g.fsg, 1, kx);
    else if (n == "dhsr") put1(g.dhsr, 1, kx);
    else if (n == "fsgr") put1(g.fsgr, 1, kx);
    else if (n == "radang") put1(g.radang, 1, il);
    else if (n == "coriol") put1(g.coriol, 1, il);
    else if (n == "sia") put1(g.sia, 1, il);
    else if (n == "coa") put1(g.coa, 1, il);
    else if (n == "cosgr") put1(g.cosgr, 1, il);
    else if (n == "cosgr2") put1(g.cosgr2, 1, il);
    else if (n == "sigl") put1(g.sigl, 1, kx);
    else if (n == "sigh") put1(g.sigh, 0, kx);
    else if (n == "grdsig") put1(g.grdsig, 1, kx);
    else if (n == "grdscp") put1(g.grdscp, 1, kx);
    else if (n == "wvi") { for (int c = 1; c <= 2; c++) for (int k = 1; k <= kx; k++) out.push_back(g.wvi[k][c]); }
    else if (n == "wt") put1(sp.wt, 1, iy);
    else if (n == "nsh2") { for (int q = 1; q <= nx; q++) out.push_back(sp.nsh2[q]); }
    else if (n == "ifac") { for (int q = 1; q <= 6; q++) out.push_back(sp.ifac[q]); }
    else if (n == "wa") put1(sp.work, 1, ix);
    else if (n == "cpol") out = sp.cpol;
    else if (n == "epsi") out = sp.epsi;
    else if (n == "el2") out = sp.el2;
    else if (n == "elm2") out = sp.elm2;
    else if (n == "el4") out = sp.el4;
    else if (n == "trfilt") out = sp.trfilt;
    else if (n == "gradx") put1(sp.gradx, 1, mx);
    else if (n == "gradym") { out = sp.gradym; for (int m = 1; m <= mx; m++) out[Spectral::i2(m, 1)] = 0.0; }
    else if (n == "gradyp") out = sp.gradyp;
    else if (n == "uvdx") out = sp.uvdx;
    else if (n == "uvdym") out = sp.uvdym;
    else if (n == "uvdyp") out = sp.uvdyp;
    else if (n == "vddym") out = sp.vddym;
    else if (n == "vddyp") out = sp.vddyp;
    else if (n == "dmp") out = im.dmp;
    else if (n == "dmpd") out = im.dmpd;
    else if (n == "dmps") out = im.dmps;
    else if (n == "dmp1") out = im.dmp1;
    else if (n == "dmp1d") out = im.dmp1d;
    else if (n == "dmp1s") out = im.dmp1s;
    else if (n == "tcorv") put1(im.tcorv, 1, kx);
    else if (n == "qcorv") put1(im.qcorv, 1, kx);
    else if (n == "tref") put1(im.tref, 1, kx);
    else if (n == "tref2") put1(im.tref2, 1, kx);
    else if (n == "tref3") put1(im.tref3, 1, kx);
    else if (n == "dhsx") put1(im.dhsx, 1, kx);
    else if (n == "elz") out = im.elz;
    else if (n == "xc") { for (int k1 = 1; k1 <= kx; k1++) for (int k = 1; k <= kx; k++) out.push_back(im.xc[k][k1]); }
    else if (n == "xd") { for (int k1 = 1; k1 <= kx; k1++) for (int k = 1; k <= kx; k++) out.push_back(im.xd[k][k1]); }
    else if (n == "xj") out = im.xj;
    else if (n == "xgeop1") out = t.var[V_xgeop1];
    else if (n == "xgeop2") out = t.var[V_xgeop2];
    else if (n == "fband") out = t.var[V_fband];
    else return -1;
    if ((int)out.size() > cap) return -(int)out.size();
    memcpy(dst, out.data(), out.size() * 8);
    return (int)out.size();
}
// implicit tables for another time step (first_step uses delt/2 and delt)
void orc_set_table_dt(double dt) { tables().imp.set_time_step(dt); }

// ---- batched per-stage entry points (Fortran-order fields, one after another) ---------------------------
void orc_rfftb(double *lines, int n) {  // raw rfftb1 on n lines of 96
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) {
        double ch[ix];
        rfftb1(ix, lines + (size_t)ix * q, ch, sp.work, sp.ifac);
    }
}
void orc_rfftf(double *lines, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) {
        double ch[ix];
        rfftf1(ix, lines + (size_t)ix * q, ch, sp.work, sp.ifac);
    }
}
void orc_legendre_inv(const double *in, double *out, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) sp.legendre_inv(in + (size_t)2 * mx * nx * q, out + (size_t)2 * mx * il * q);
}
void orc_legendre_dir(const double *in, double *out, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) sp.legendre_dir(in + (size_t)2 * mx * il * q, out + (size_t)2 * mx * nx * q);
}
void orc_fourier_inv(const double *in, double *out, int kcos, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) sp.fourier_inv(in + (size_t)2 * mx * il * q, out + (size_t)ix * il * q, kcos);
}
void orc_fourier_dir(const double *in, double *out, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++) sp.fourier_dir(in + (size_t)ix * il * q, out + (size_t)2 * mx * il * q);
}
void orc_spec2grid(const double *in, double *out, int kcos, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++)
        sp.spec2grid(S2{(cplx *)in + (size_t)mx * nx * q, mx}, G2{out + (size_t)ix * il * q, ix}, kcos);
}
void orc_grid2spec(const double *in, double *out, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++)
        sp.grid2spec(G2{(double *)in + (size_t)ix * il * q, ix}, S2{(cplx *)out + (size_t)mx * nx * q, mx});
}
#define SPEC(p, q) S2{(cplx *)(p) + (size_t)mx * nx * (q), mx}
void orc_vort2vel(const double *vor, const double *div, double *u, double *v, int n) {
    const Spectral &sp = tables().spec;
    for (int q = 0; q < n; q++) sp.vort2vel(SPEC(vor, q), SPEC(div, q), SPEC(u, q), SPEC(v, q));
}
void orc_vel2vort(const double *u, const double *v, double *vor, double *div, int n) {
    const Spectral &sp = tables().spec;
    for (int q = 0; q < n; q++) sp.vel2vort(SPEC(u, q), SPEC(v, q), SPEC(vor, q), SPEC(div, q));
}
void orc_gradient(const double *psi, double *dx, double *dy, int n) {
    const Spectral &sp = tables().spec;
    for (int q = 0; q < n; q++) sp.gradient(SPEC(psi, q), SPEC(dx, q), SPEC(dy, q));
}
void orc_laplacian(const double *in, double *out, int inverse, int n) {
    const Spectral &sp = tables().spec;
    for (int q = 0; q < n; q++) {
        if (inverse) sp.laplacian_inv(SPEC(in, q), SPEC(out, q)); else sp.laplacian(SPEC(in, q), SPEC(out, q));
    }
}
void orc_grid_vel2vort(const double *ug, const double *vg, double *vor, double *div, int kcos, int n) {
    const Spectral &sp = tables().spec;
#pragma omp parallel for
    for (int q = 0; q < n; q++)
        sp.grid_vel2vort(G2{(double *)ug + (size_t)ix * il * q, ix}, G2{(double *)vg + (size_t)ix * il * q, ix},
                         SPEC(vor, q), SPEC(div, q), kcos);
}

// column physics on explicit grid inputs (physics.f90:103-231); state supplies surface/forcing fields and
// receives the diagnostics.  Tendency arrays are in/out (dynamics tendencies + physics).
void orc_physics_columns(void *st, const double *ug, const double *vg, const double *tg, double *qg,
                         const double *phig, const double *pslg, double *utend, double *vtend, double *ttend,
                         double *qtend, int *dbg) {
    State &s = *(State *)st;
    physics_columns(s, G3{(double *)ug, ix, il}, G3{(double *)vg, ix, il}, G3{(double *)tg, ix, il}, G3{qg, ix, il},
                    G3{(double *)phig, ix, il}, G2{(double *)pslg, ix}, G3{utend, ix, il}, G3{vtend, ix, il},
                    G3{ttend, ix, il}, G3{qtend, ix, il}, dbg);
}
// prepares a bare State so that physics_columns can run on it: tables only
void orc_state_init_tables(void *st) {
    State &s = *(State *)st;
    s.geo.initialize();
    s.spec.initialize(&s.geo);
    s.imp.initialize(&s.geo);
    s.imp.set_time_step(2 * delt);
    initialize_geopotential(s);
    radset(s.p(V_fband));
}
void orc_zonal_average_fields(void *st, double tyear) { get_zonal_average_fields(*(State *)st, tyear); }

// full tendencies of one state (tendencies.f90:11-39), outputs (mx,nx,kx) complex each, psdt (mx,nx)
void orc_tendencies(void *st, int j2, double *vordt, double *divdt, double *tdt, double *psdt, double *trdt) {
    State &s = *(State *)st;
    get_tendencies(s, S3{(cplx *)vordt, mx, nx}, S3{(cplx *)divdt, mx, nx}, S3{(cplx *)tdt, mx, nx},
                   S2{(cplx *)psdt, mx}, S3{(cplx *)trdt, mx, nx}, j2);
}
void orc_raw_step(void *st, int j1, int j2, double dt) { step(*(State *)st, j1, j2, dt); }
// implicit_terms (implicit.f90:234-289) in place on (mx,nx,kx) divdt, tdt and (mx,nx) psdt, with the state's time step
void orc_implicit_terms(void *st, double *divdt, double *tdt, double *psdt) {
    ((State *)st)->imp.implicit_terms(S3{(cplx *)divdt, mx, nx}, S3{(cplx *)tdt, mx, nx}, S2{(cplx *)psdt, mx});
}
// horizontal diffusion + time integration of step() for GIVEN tendencies (time_stepping.f90:78-144); tendencies are clobbered
void orc_apply_tendencies(void *st, int j1, double dt, double *vordt, double *divdt, double *tdt, double *psdt, double *trdt) {
    apply_tendencies(*(State *)st, j1, dt, S3{(cplx *)vordt, mx, nx}, S3{(cplx *)divdt, mx, nx}, S3{(cplx *)tdt, mx, nx},
                     S2{(cplx *)psdt, mx}, S3{(cplx *)trdt, mx, nx});
}
void orc_set_time_step(void *st, double dt) { ((State *)st)->imp.set_time_step(dt); }
void orc_set_forcing(void *st, void *ctl, int imode) {
    Control &c = *(Control *)ctl;
    set_forcing(*(State *)st, imode, c.model_datetime, c.tyear);
}
void orc_couple(void *st, void *ctl, int day) {
    Control &c = *(Control *)ctl;
    couple_land_atm(*(State *)st, day, c.imont1, c.tmonth);
    couple_sea_atm(*(State *)st, day, c);
}
void orc_advance_date(void *ctl) { ((Control *)ctl)->advance_date(); }
// SPPT switch of one member (sppt.f90; see gen_sppt): counter-based generator keyed by (seed, member id)
void orc_set_sppt(void *st, int on, unsigned long long seed, unsigned long long member) {
    State &s = *(State *)st;
    s.sppt_on = on != 0, s.sppt_seed = seed, s.sppt_member = member, s.sppt_calls = 0;
    s.sppt_spec.clear();
}
// the member's AR(1) pattern in spectral space (mx,nx,kx complex) and the grid-point pattern of the last step (ix,il,kx)
int orc_get_sppt(void *st, double *spec, double *grid) {
    State &s = *(State *)st;
    if (s.sppt_spec.empty() || s.sppt_last.empty()) return -1;
    memcpy(spec, s.sppt_spec.data(), sizeof(cplx) * s.sppt_spec.size());
    memcpy(grid, s.sppt_last.data(), sizeof(double) * s.sppt_last.size());
    return 0;
}

}  // extern "C"
