// ORACLE (test infrastructure) -- model state, calendar, boundary/land/sea models, daily forcing,
// initialisation and the single-step driver.
// Follows model_state.f90, model_control.f90, boundaries.f90, interpolation.f90, land_model.f90, sea_model.f90,
// coupler.f90, forcing.f90, prognostics.f90, initialization.f90 and speedy.f90 of the reference.
#include "speedy_oracle.hpp"

namespace orc {

static const size_t NG = (size_t)ix * il;
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------------
// model_state.f90:358-... (ModelState_allocate): every array allocated and zeroed
State::State() {
    for (int v = 0; v < SPDY_NVARS; v++) {
        const spdy_vardef &d = SPDY_VARDEFS[v];
        if (d.ndim == 0 || d.kind == SPDY_F4) continue;
        size_t n = 1;
        bool dyn = false;
        for (int q = 0; q < d.ndim; q++) {
            if (d.dims[q] < 0) dyn = true;
            n *= (size_t)(d.dims[q] < 0 ? 1 : d.dims[q]);
        }
        if (dyn) continue;  // sst_anom: allocated by alloc_sst_anom
        if (d.kind == SPDY_C16) n *= 2;
        var[v].assign(n, 0.0);
    }
    lon.assign(ix, 0.f), lat.assign(il, 0.f), lev.assign(kx, 0.f);
}
void State::alloc_sst_anom(int n_months_) {  // speedy_driver.f90.j2 modelstate_init_sst_anom: (ix,il,0:n_months+1)
    n_months = n_months_;
    var[V_sst_anom].assign(NG * (size_t)(n_months + 2), 0.0);
}

// ---------------------------------------------------------------------------------------------------
// model_control.f90:73-186
static const int ncal = 365;
static const int ncal365[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
void Control::update_forcing_params() {
    imont1 = model_datetime.month;
    // REAL(4) arithmetic: (day - 0.5)/float(n)
    tmonth = (double)(((float)model_datetime.day - 0.5f) / (float)ndaycal[model_datetime.month][1]);
    tyear = (double)(((float)(ndaycal[model_datetime.month][2] + model_datetime.day) - 0.5f) / (float)ncal);
}
void Control::initialize(const Datetime &s, const Datetime &e) {
    start_datetime = s;
    end_datetime = e;
    model_datetime = s;
    for (int jm = 1; jm <= 12; jm++) ndaycal[jm][1] = ncal365[jm - 1];
    ndaycal[1][2] = 0;
    for (int jm = 2; jm <= 12; jm++) ndaycal[jm][2] = ndaycal[jm - 1][1] + ndaycal[jm - 1][2];
    month_idx = 1;
    update_forcing_params();
}
void Control::advance_date() {
    Datetime &d = model_datetime;
    d.minute += 24 * 60 / nsteps;
    if (d.minute >= 60) {
        d.minute %= 60;
        d.hour += 1;
    }
    if (d.hour >= 24) {
        d.hour %= 24;
        d.day += 1;
    }
    if (d.year % 4 == 0 && d.month == 2) {
        if (d.day > 29) {
            d.day = 1;
            d.month += 1;
            month_idx += 1;
        }
    } else if (d.day > ndaycal[d.month][1]) {
        d.day = 1;
        d.month += 1;
        month_idx += 1;
    }
    if (d.month > 12) {
        d.month = 1;
        d.year += 1;
    }
    update_forcing_params();
}

// ---------------------------------------------------------------------------------------------------
// boundaries.f90:22-114
void initialize_boundaries(State &s) {
    G2 phi0 = s.g2(V_phi0), orog = s.g2(V_orog);
    for (size_t q = 0; q < NG; q++) phi0.p[q] = grav * orog.p[q];
    s.spec.grid_filter(phi0, s.g2(V_phis0));
}
static void check_surface_fields(G2 fmask, int nf, double fmin, double fmax, double fset, double *field) {
    (void)fmin, (void)fmax;  // the reference only counts out-of-range points (nfault), it never uses the count
    for (int jf = 0; jf < nf; jf++)
        for (size_t q = 0; q < NG; q++)
            if (!(fmask.p[q] > 0.0)) field[q + NG * jf] = fset;
}
static void fill_missing_values(double *sf /*(ix,il)*/, double fmis) {
    double sf2[ix + 2];
    static double fmean = 0.0;  // implicit SAVE in the reference (boundaries.f90:77); init path only
    int j1 = 0, j2 = 0, j3 = 0;
    for (int hemisphere = 1; hemisphere <= 2; hemisphere++) {
        if (hemisphere == 1) {
            j1 = il / 2, j2 = 1, j3 = -1;
        } else {
            j1 = j1 + 1, j2 = il, j3 = 1;
        }
        for (int j = j1; (j3 > 0) ? (j <= j2) : (j >= j2); j += j3) {
            int nmis = 0;
            for (int i = 1; i <= ix; i++) sf2[i] = sf[(i - 1) + ix * (j - 1)];
            for (int i = 1; i <= ix; i++)
                if (sf[(i - 1) + ix * (j - 1)] < fmis) {
                    nmis++;
                    sf2[i] = 0.0;
                }
            if (nmis < ix) {
                double sum = 0.0;
                for (int i = 1; i <= ix; i++) sum += sf2[i];
                fmean = sum / (double)(float)(ix - nmis);
            }
            for (int i = 1; i <= ix; i++)
                if (sf[(i - 1) + ix * (j - 1)] < fmis) sf2[i] = fmean;
            sf2[0] = sf2[ix];
            sf2[ix + 1] = sf2[1];
            for (int i = 1; i <= ix; i++)
                if (sf[(i - 1) + ix * (j - 1)] < fmis) sf[(i - 1) + ix * (j - 1)] = 0.5 * (sf2[i - 1] + sf2[i + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// interpolation.f90:17-94
static void monthly_interp(int month_idx, const double *in_field /*(ix,il,0:n-1)*/, double *out, double mf) {
    int imon2;
    double wmon;
    if (mf <= 0.5) {
        imon2 = month_idx - 1;
        wmon = 0.5 - mf;
    } else {
        imon2 = month_idx + 1;
        wmon = mf - 0.5;
    }
    for (size_t q = 0; q < NG; q++)
        out[q] = in_field[q + NG * month_idx] + wmon * (in_field[q + NG * imon2] - in_field[q + NG * month_idx]);
}
static void forint(int imon, const double *for12, double *for1, double tmonth) {
    int imon2;
    double wmon;
    if (tmonth <= 0.5) {
        imon2 = imon - 1;
        if (imon == 1) imon2 = 12;
        wmon = 0.5 - tmonth;
    } else {
        imon2 = imon + 1;
        if (imon == 12) imon2 = 1;
        wmon = tmonth - 0.5;
    }
    for (size_t q = 0; q < NG; q++)
        for1[q] = for12[q + NG * (imon - 1)] + wmon * (for12[q + NG * (imon2 - 1)] - for12[q + NG * (imon - 1)]);
}
static void forin5(int imon, const double *for12, double *for1, double tmonth) {
    int im2 = imon - 2, im1 = imon - 1, ip1 = imon + 1, ip2 = imon + 2;
    if (im2 < 1) im2 += 12;
    if (im1 < 1) im1 += 12;
    if (ip1 > 12) ip1 -= 12;
    if (ip2 > 12) ip2 -= 12;
    const double c0 = (double)(1.0f / 12.0f);
    const double t0 = c0 * tmonth, t1 = c0 * (1.0 - tmonth), t2 = 0.25 * tmonth * (1 - tmonth);
    const double wm2 = -t1 + t2, wm1 = -c0 + 8 * t1 - 6 * t2, w0 = 7 * c0 + 10 * t2, wp1 = -c0 + 8 * t0 - 6 * t2,
                 wp2 = -t0 + t2;
    for (size_t q = 0; q < NG; q++)
        for1[q] = wm2 * for12[q + NG * (im2 - 1)] + wm1 * for12[q + NG * (im1 - 1)] + w0 * for12[q + NG * (imon - 1)] +
                  wp1 * for12[q + NG * (ip1 - 1)] + wp2 * for12[q + NG * (ip2 - 1)];
}

// ---------------------------------------------------------------------------------------------------
// land_model.f90:18-148
static const double snow_depth2cover = FL(60.0);
void land_model_init(State &s) {
    const double swcap = FL(0.30), swwil = FL(0.17), thrsh = FL(0.1);
    G2 fml = s.g2(V_fmask_land), fmo = s.g2(V_fmask_orig), bml = s.g2(V_bmask_land);
    Grid2 veg, dmask;
    for (size_t q = 0; q < NG; q++) fml.p[q] = fmo.p[q];
    for (size_t q = 0; q < NG; q++) {
        if (fml.p[q] >= thrsh) {
            bml.p[q] = 1.0;
            if (fmo.p[q] > (1.0 - thrsh)) fml.p[q] = 1.0;
        } else {
            bml.p[q] = 0.0;
            fml.p[q] = 0.0;
        }
    }
    for (int month = 1; month <= 12; month++) fill_missing_values(s.p(V_stl12) + NG * (month - 1), 0.0);
    check_surface_fields(bml, 12, 0.0, 400.0, 273.0, s.p(V_stl12));
    check_surface_fields(bml, 12, 0.0, 20000.0, 0.0, s.p(V_snowd12));
    for (size_t q = 0; q < NG; q++) veg.d[q] = dmax(0.0, s.p(V_veg_high)[q] + FL(0.8) * s.p(V_veg_low)[q]);
    const int idep2 = 3;
    const double swwil2 = idep2 * swwil;
    const double rsw = 1.0 / (swcap + idep2 * (swcap - swwil));
    for (int month = 1; month <= 12; month++)
        for (size_t q = 0; q < NG; q++) {
            double swroot = idep2 * s.p(V_soil_wc_l2)[q + NG * (month - 1)];
            s.p(V_soilw12)[q + NG * (month - 1)] =
                dmin(1.0, rsw * (s.p(V_soil_wc_l1)[q + NG * (month - 1)] + veg.d[q] * dmax(0.0, swroot - swwil2)));
        }
    check_surface_fields(bml, 12, 0.0, 10.0, 0.0, s.p(V_soilw12));
    const double depth_soil = 1.0, depth_lice = 5.0, tdland = 40.0;
    const double flandmin = (double)(1.0f / 3.0f);
    const double hcapl = depth_soil * FL(2.50e+6), hcapli = depth_lice * FL(1.93e+6);
    for (size_t q = 0; q < NG; q++) {
        dmask.d[q] = 1.0;
        if (fml.p[q] < flandmin) dmask.d[q] = 0.0;
    }
    for (size_t q = 0; q < NG; q++) s.p(V_rhcapl)[q] = (s.p(V_alb0)[q] < FL(0.4)) ? delt / hcapl : delt / hcapli;
    for (size_t q = 0; q < NG; q++) s.p(V_cdland)[q] = dmask.d[q] * tdland / (1.0 + dmask.d[q] * tdland);
}

// land_model.f90:196-215
static void run_land_model(State &s) {
    double *stl_lm = s.p(V_stl_lm), *stlcl = s.p(V_stlcl_obs), *cdland = s.p(V_cdland), *rhcapl = s.p(V_rhcapl);
    double *hfluxn1 = s.p(V_hfluxn);
    for (size_t q = 0; q < NG; q++) {
        double tanom = stl_lm[q] - stlcl[q];
        tanom = cdland[q] * (tanom + rhcapl[q] * hfluxn1[q]);
        stl_lm[q] = tanom + stlcl[q];
    }
}

// land_model.f90:151-192
void couple_land_atm(State &s, int day, int imont1, double tmonth) {
    forin5(imont1, s.p(V_stl12), s.p(V_stlcl_obs), tmonth);
    forint(imont1, s.p(V_snowd12), s.p(V_snowdcl_obs), tmonth);
    forint(imont1, s.p(V_soilw12), s.p(V_soilwcl_obs), tmonth);
    if (day == 0) {
        for (size_t q = 0; q < NG; q++) s.p(V_stl_lm)[q] = s.p(V_stlcl_obs)[q], s.p(V_land_temp)[q] = s.p(V_stlcl_obs)[q];
    } else if (s.land_coupling_flag) {
        run_land_model(s);
        for (size_t q = 0; q < NG; q++) s.p(V_land_temp)[q] = s.p(V_stl_lm)[q];
    } else {
        for (size_t q = 0; q < NG; q++) s.p(V_land_temp)[q] = s.p(V_stlcl_obs)[q];
    }
    for (size_t q = 0; q < NG; q++) {
        s.p(V_snow_depth)[q] = s.p(V_snowdcl_obs)[q];
        s.p(V_soil_avail_water)[q] = s.p(V_soilwcl_obs)[q];
    }
}

// ---------------------------------------------------------------------------------------------------
// sea_model.f90:33-191   (sea_coupling_flag = 0, ice_coupling_flag = 1: sea_model.f90:20,23)
static const double beta = FL(1.0);
void sea_model_init(State &s) {
    const Geometry &g = s.geo;
    const double depth_ml = FL(60.), dept0_ml = FL(40.), depth_ice = FL(2.5), dept0_ice = FL(1.5), tdsst = FL(90.),
                 fseamin = (double)(1.0f / 3.0f), tdice = FL(30.0), thrsh = FL(0.1);
    double hcaps[il + 1], hcapi[il + 1];
    Grid2 dmask;
    G2 fms = s.g2(V_fmask_sea), fmo = s.g2(V_fmask_orig), bms = s.g2(V_bmask_sea);
    for (size_t q = 0; q < NG; q++) {
        fms.p[q] = 1.0 - fmo.p[q];
        if (fms.p[q] >= thrsh) {
            bms.p[q] = 1.0;
            if (fms.p[q] > (1.0 - thrsh)) fms.p[q] = 1.0;
        } else {
            bms.p[q] = 0.0;
            fms.p[q] = 0.0;
        }
    }
    double *deglat_s = s.p(V_deglat_s) - 1;
    for (int j = 1; j <= il; j++) deglat_s[j] = g.radang[j] * FL(90.0) / F_ASIN1;
    for (int month = 1; month <= 12; month++) fill_missing_values(s.p(V_sst12) + NG * (month - 1), 0.0);
    check_surface_fields(bms, 12, 100.0, 400.0, 273.0, s.p(V_sst12));
    for (size_t q = 0; q < NG * 12; q++) s.p(V_sea_ice_frac12)[q] = dmax(s.p(V_sea_ice_frac12)[q], 0.0);
    check_surface_fields(bms, 12, 0.0, 1.0, 0.0, s.p(V_sea_ice_frac12));
    if (!s.var[V_sst_anom].empty()) check_surface_fields(bms, 3, -50.0, 50.0, 0.0, s.p(V_sst_anom));
    for (size_t q = 0; q < NG; q++) s.p(V_hfseacl)[q] = 0.0;
    const double crad = (double)((float)F_ASIN1 / 90.0f);  // asin(1.)/90. in REAL(4)
    for (int j = 1; j <= il; j++) {
        double coslat = cos(crad * deglat_s[j]);
        hcaps[j] = FL(4.18e+6) * (depth_ml + (dept0_ml - depth_ml) * (coslat * coslat * coslat));
        hcapi[j] = FL(1.93e+6) * (depth_ice + (dept0_ice - depth_ice) * (coslat * coslat));
    }
    for (size_t q = 0; q < NG; q++) dmask.d[q] = 1.0;  // l_globe
    G2 rhcaps = s.g2(V_rhcaps), rhcapi = s.g2(V_rhcapi);
    for (int j = 2; j <= il - 1; j++)
        for (int i = 1; i <= ix; i++) rhcaps(i, j) = 0.25 * (dmask(i, j - 1) + 2 * dmask(i, j) + dmask(i, j + 1));
    for (int j = 2; j <= il - 1; j++)
        for (int i = 1; i <= ix; i++) dmask(i, j) = rhcaps(i, j);
    for (size_t q = 0; q < NG; q++)
        if (fms.p[q] < fseamin) dmask.d[q] = 0.0;
    for (int j = 1; j <= il; j++)
        for (int i = 1; i <= ix; i++) {
            rhcaps(i, j) = delt / hcaps[j];
            rhcapi(i, j) = delt / hcapi[j];
        }
    for (size_t q = 0; q < NG; q++) {
        s.p(V_cdsea)[q] = dmask.d[q] * tdsst / (1.0 + dmask.d[q] * tdsst);
        s.p(V_cdice)[q] = dmask.d[q] * tdice / (1.0 + dmask.d[q] * tdice);
    }
}

// sea_model.f90:313-383
static void run_sea_model(State &s) {
    const double sstfr = (double)(273.2f - 1.8f);
    const double anom0 = 20.0;
    double *ssrd = s.p(V_ssrd), *tice_am = s.p(V_tice_am), *shf2 = s.p(V_shf) + NG, *evap2 = s.p(V_evap) + NG,
           *hfluxn2 = s.p(V_hfluxn) + NG, *sice_am = s.p(V_sice_am), *hfseacl = s.p(V_hfseacl),
           *sicecl_ob = s.p(V_sicecl_ob), *tice_om = s.p(V_tice_om), *sst_om = s.p(V_sst_om),
           *sstcl_ob = s.p(V_sstcl_ob), *cdsea = s.p(V_cdsea), *rhcaps = s.p(V_rhcaps), *ticecl_ob = s.p(V_ticecl_ob),
           *cdice = s.p(V_cdice), *rhcapi = s.p(V_rhcapi), *sice_om = s.p(V_sice_om);
    for (size_t q = 0; q < NG; q++) {
        double difice = (albsea - albice) * ssrd[q] + emisfc * sbc * (pow(sstfr, 4.0) - pow(tice_am[q], 4.0)) + shf2[q] +
                        evap2[q] * alhc;
        double hflux_i = hfluxn2[q] + difice * (1.0 - sice_am[q]);
        double hflux = hfluxn2[q] - hfseacl[q] - sicecl_ob[q] * (hflux_i + beta * (sstfr - tice_om[q]));
        double tanom = sst_om[q] - sstcl_ob[q];
        tanom = cdsea[q] * (tanom + rhcaps[q] * hflux);
        sst_om[q] = tanom + sstcl_ob[q];
        hflux = hflux_i + beta * (sstfr - tice_om[q]);
        tanom = tice_om[q] - ticecl_ob[q];
        double cdis = cdice[q] * (anom0 / (anom0 + fabs(tanom)));
        tanom = cdis * (tanom + rhcapi[q] * hflux);
        tice_om[q] = tanom + ticecl_ob[q];
        sice_om[q] = sicecl_ob[q];
    }
}

// sea_model.f90:193-310
void couple_sea_atm(State &s, int day, const Control &c) {
    double *sstcl_ob = s.p(V_sstcl_ob), *sicecl_ob = s.p(V_sicecl_ob), *ticecl_ob = s.p(V_ticecl_ob);
    forin5(c.imont1, s.p(V_sst12), sstcl_ob, c.tmonth);
    forint(c.imont1, s.p(V_sea_ice_frac12), sicecl_ob, c.tmonth);
    if (s.sst_anomaly_coupling_flag && !s.var[V_sst_anom].empty())
        monthly_interp(c.month_idx, s.p(V_sst_anom), s.p(V_sstan_ob), c.tmonth);
    const double sstfr = (double)(273.2f - 1.8f);
    for (size_t q = 0; q < NG; q++) {
        if (sstcl_ob[q] > sstfr) {
            sicecl_ob[q] = dmin(0.5, sicecl_ob[q]);
            ticecl_ob[q] = sstfr;
            if (sicecl_ob[q] > 0.0) sstcl_ob[q] = sstfr + (sstcl_ob[q] - sstfr) / (1.0 - sicecl_ob[q]);
        } else {
            sicecl_ob[q] = dmax(0.5, sicecl_ob[q]);
            ticecl_ob[q] = sstfr + (sstcl_ob[q] - sstfr) / sicecl_ob[q];
            sstcl_ob[q] = sstfr;
        }
    }
    if (day == 0) {
        for (size_t q = 0; q < NG; q++) {
            s.p(V_sst_om)[q] = sstcl_ob[q];
            s.p(V_tice_om)[q] = ticecl_ob[q];
            s.p(V_sice_om)[q] = sicecl_ob[q];
            s.p(V_sst_om)[q] = 0.0;  // sea_coupling_flag <= 0
            s.p(V_wsst_ob)[q] = 0.0;
        }
    } else {
        run_sea_model(s);  // ice_coupling_flag > 0
    }
    for (size_t q = 0; q < NG; q++) {
        s.p(V_sstan_am)[q] = 0.0;
        if (s.sst_anomaly_coupling_flag) s.p(V_sstan_am)[q] = s.p(V_sstan_ob)[q];
        s.p(V_sst_am)[q] = sstcl_ob[q] + s.p(V_sstan_am)[q];
        s.p(V_sice_am)[q] = s.p(V_sice_om)[q];
        s.p(V_tice_am)[q] = s.p(V_tice_om)[q];
        s.p(V_sst_am)[q] = s.p(V_sst_am)[q] + s.p(V_sice_am)[q] * (s.p(V_tice_am)[q] - s.p(V_sst_am)[q]);
        s.p(V_ssti_om)[q] = s.p(V_sst_om)[q] + s.p(V_sice_am)[q] * (s.p(V_tice_am)[q] - s.p(V_sst_om)[q]);
    }
}

// ---------------------------------------------------------------------------------------------------
// forcing.f90:15-117
void set_forcing(State &s, int imode, const Datetime &dt, double tyear) {
    Grid2 corh, tsfc, tref, psfc, qsfc, qref, ones;
    double gamlat[il + 1];
    if (imode == 0) {
        radset(s.p(V_fband));
        set_orog_land_sfc_drag(s.g2(V_phis0), s.g2(V_forog));
        s.ablco2_ref = s.air_absortivity_co2;
    }
    get_zonal_average_fields(s, tyear);
    for (size_t q = 0; q < NG; q++) {
        s.p(V_snowc)[q] = dmin(1.0, s.p(V_snow_depth)[q] / snow_depth2cover);
        s.p(V_alb_land)[q] = s.p(V_alb0)[q] + s.p(V_snowc)[q] * (albsn - s.p(V_alb0)[q]);
        s.p(V_alb_sea)[q] = albsea + s.p(V_sice_am)[q] * (albice - albsea);
        s.p(V_alb_surface)[q] = s.p(V_alb_sea)[q] + s.p(V_fmask_land)[q] * (s.p(V_alb_land)[q] - s.p(V_alb_sea)[q]);
    }
    if (s.increase_co2) {
        const int iyear_ref = 1950;
        const double del_co2 = FL(0.005);
        s.air_absortivity_co2 = s.ablco2_ref * exp(del_co2 * ((double)dt.year + tyear - (double)iyear_ref));
    }
    gamlat[1] = gamma_ / (FL(1000.) * grav);
    for (int j = 2; j <= il; j++) gamlat[j] = gamlat[1];
    G2 phis0 = s.g2(V_phis0);
    for (int j = 1; j <= il; j++)
        for (int i = 1; i <= ix; i++) corh(i, j) = gamlat[j] * phis0(i, j);
    s.spec.grid2spec(corh, S2{s.imp.tcorh.data(), mx});
    for (int j = 1; j <= il; j++) {
        double pexp = 1.0 / (rgas * gamlat[j]);
        for (int i = 1; i <= ix; i++) {
            size_t q = (i - 1) + (size_t)ix * (j - 1);
            tsfc.d[q] = s.p(V_fmask_land)[q] * s.p(V_land_temp)[q] + s.p(V_fmask_sea)[q] * s.p(V_sst_am)[q];
            tref.d[q] = tsfc.d[q] + corh.d[q];
            psfc.d[q] = pow(tsfc.d[q] / tref.d[q], pexp);
        }
    }
    for (size_t q = 0; q < NG; q++) ones.d[q] = psfc.d[q] / psfc.d[q];
    get_qsat(tref.d.data(), ones.d.data(), -1.0, qref.d.data(), (int)NG);
    get_qsat(tsfc.d.data(), psfc.d.data(), 1.0, qsfc.d.data(), (int)NG);
    for (size_t q = 0; q < NG; q++) corh.d[q] = refrh1 * (qref.d[q] - qsfc.d[q]);
    s.spec.grid2spec(corh, S2{s.imp.qcorh.data(), mx});
}

// ---------------------------------------------------------------------------------------------------
// prognostics.f90:40-117
int initialize_prognostics(State &s) {
    const Spectral &sp = s.spec;
    const Geometry &g = s.geo;
    Spec2 surfs;
    Grid2 surfg;
    const double gam1 = gamma_ / (FL(1000.0) * grav);
    S2 phis = s.s2(V_phis);
    sp.grid2spec(s.g2(V_phis0), phis);
    S3 vor1 = s.s4lev(V_vor, 1), div1 = s.s4lev(V_div, 1), t1 = s.s4lev(V_t, 1), tr1 = s.s4lev(V_tr, 1);
    const size_t ns = (size_t)mx * nx;
    for (size_t q = 0; q < ns * kx; q++) vor1.p[q] = div1.p[q] = tr1.p[q] = cplx{0.0, 0.0};
    const double tref = 288.0, ttop = 216.0;
    const double gam2 = gam1 / tref, rgam = rgas * gam1, rgamr = 1.0 / rgam;
    for (size_t q = 0; q < ns * 2; q++) t1.p[q] = cplx{0.0, 0.0};
    for (size_t q = 0; q < ns; q++) surfs.d[q] = (-gam1) * phis.p[q];
    // sqrt(2.0)*(1.0,0.0)*ttop : REAL(4) sqrt(2.) times COMPLEX(4) (1,0), then times REAL(8)
    t1(1, 1, 1) = cplx{F_SQRT2 * ttop, 0.0 * ttop};
    t1(1, 1, 2) = cplx{F_SQRT2 * ttop, 0.0 * ttop};
    surfs(1, 1) = cplx{F_SQRT2 * tref, 0.0 * tref} - gam1 * phis(1, 1);
    for (int k = 3; k <= kx; k++) {
        double f = pow(g.fsg[k], rgam);
        for (size_t q = 0; q < ns; q++) t1.p[q + ns * (k - 1)] = surfs.d[q] * f;
    }
    const double rlog0 = F_LOG1013;
    G2 phis0 = s.g2(V_phis0);
    for (size_t q = 0; q < NG; q++) surfg.d[q] = rlog0 + rgamr * log(1.0 - gam2 * phis0.p[q]);
    S2 ps1 = s.s3lev(V_ps, 1);
    sp.grid2spec(surfg, ps1);
    sp.truncate(ps1);
    const double esref = 17.0;
    const double qref = refrh1 * FL(0.622) * esref;
    const double qexp = hscale / hshum;
    for (size_t q = 0; q < NG; q++) surfg.d[q] = qref * exp(qexp * surfg.d[q]);
    sp.grid2spec(surfg, surfs);
    sp.truncate(surfs);
    for (int k = 3; k <= kx; k++) {
        double f = pow(g.fsg[k], qexp);
        for (size_t q = 0; q < ns; q++) tr1.p[q + ns * (k - 1)] = surfs.d[q] * f;
    }
    return check_diagnostics(s, 1);
}

// prognostics.f90:125-154
void spectral2grid(State &s) {
    const Spectral &sp = s.spec;
    Spec2 ucos, vcos;
    S3 vor = s.s4lev(V_vor, 1), div = s.s4lev(V_div, 1), t = s.s4lev(V_t, 1), tr = s.s4lev(V_tr, 1), phi = s.s3(V_phi);
    G3 ug = s.g3(V_u_grid), vg = s.g3(V_v_grid), tg = s.g3(V_t_grid), qg = s.g3(V_q_grid), pg = s.g3(V_phi_grid);
    for (int k = 1; k <= kx; k++) {
        sp.vort2vel(vor.slab(k), div.slab(k), ucos, vcos);
        sp.spec2grid(ucos, ug.slab(k), 2);
        sp.spec2grid(vcos, vg.slab(k), 2);
        sp.spec2grid(t.slab(k), tg.slab(k), 1);
        sp.spec2grid(tr.slab(k), qg.slab(k), 1);
        sp.spec2grid(phi.slab(k), pg.slab(k), 1);
        for (size_t q = 0; q < NG; q++) {
            qg.p[q + NG * (k - 1)] = qg.p[q + NG * (k - 1)] * FL(1.0e-3);
            pg.p[q + NG * (k - 1)] = pg.p[q + NG * (k - 1)] / grav;
        }
    }
    G2 psg = s.g2(V_ps_grid);
    sp.spec2grid(s.s3lev(V_ps, 1), psg, 1);
    for (size_t q = 0; q < NG; q++) psg.p[q] = p0 * exp(psg.p[q]);
}

// prognostics.f90:157-176
void grid2spectral(State &s) {
    const Spectral &sp = s.spec;
    S3 vor = s.s4lev(V_vor, 1), div = s.s4lev(V_div, 1), t = s.s4lev(V_t, 1), tr = s.s4lev(V_tr, 1), phi = s.s3(V_phi);
    G3 ug = s.g3(V_u_grid), vg = s.g3(V_v_grid), tg = s.g3(V_t_grid), qg = s.g3(V_q_grid), pg = s.g3(V_phi_grid);
    const size_t ns = (size_t)mx * nx;
    for (int k = 1; k <= kx; k++) {
        sp.grid_vel2vort(ug.slab(k), vg.slab(k), vor.slab(k), div.slab(k), 2);
        sp.grid2spec(tg.slab(k), t.slab(k));
        sp.grid2spec(qg.slab(k), tr.slab(k));
        sp.grid2spec(pg.slab(k), phi.slab(k));
        for (size_t q = 0; q < ns; q++) {
            cplx &a = tr.p[q + ns * (k - 1)], &b = phi.p[q + ns * (k - 1)];
            a = cplx{a.re / FL(1.0e-3), a.im / FL(1.0e-3)};
            b = b * grav;
        }
    }
    Grid2 tmp;
    G2 psg = s.g2(V_ps_grid);
    for (size_t q = 0; q < NG; q++) tmp.d[q] = log(psg.p[q] / p0);
    sp.grid2spec(tmp, s.s3lev(V_ps, 1));
}

// prognostics.f90:180-219
void grid_filter_state(State &s) {
    const Spectral &sp = s.spec;
    Grid2 tmp;
    const int ids[5] = {V_u_grid, V_v_grid, V_t_grid, V_q_grid, V_phi_grid};
    for (int k = 1; k <= kx; k++)
        for (int v = 0; v < 5; v++) {
            G2 f = s.g3(ids[v]).slab(k);
            sp.grid_filter(f, tmp);
            for (size_t q = 0; q < NG; q++) f.p[q] = tmp.d[q];
        }
    G2 f = s.g2(V_ps_grid);
    sp.grid_filter(f, tmp);
    for (size_t q = 0; q < NG; q++) f.p[q] = tmp.d[q];
}

// ---------------------------------------------------------------------------------------------------
// initialization.f90:13-91
int initialize_state(State &s, Control &c) {
    s.current_step = 0;
    s.geo.initialize();
    s.spec.initialize(&s.geo);
    s.imp.initialize(&s.geo);
    initialize_geopotential(s);
    initialize_boundaries(s);
    int err = initialize_prognostics(s);
    if (err != 0) return err;
    // coupler.f90:13-30
    land_model_init(s);
    couple_land_atm(s, 0, c.imont1, c.tmonth);
    sea_model_init(s);
    couple_sea_atm(s, 0, c);
    set_forcing(s, 0, c.model_datetime, c.tyear);
    first_step(s);
    for (int k = 1; k <= kx; k++) s.lev[k - 1] = (float)s.geo.fsg[k];
    for (int k = 0; k < ix; k++) s.lon[k] = 3.75f * (float)k;
    for (int k = 1; k <= il; k++) s.lat[k - 1] = (float)s.geo.radang[k] * 90.0f / (float)F_ASIN1;
    s.initialized = true;
    return 0;
}

// speedy.f90:20-74
int do_single_step(State &s, Control &c) {
    if (!s.initialized) return -1;
    if (s.current_step % nsteps == 0) set_forcing(s, 1, c.model_datetime, c.tyear);
    s.compute_shortwave = (s.current_step % nstrad == 0);
    step(s, 2, 2, 2 * delt);
    s.current_step += 1;
    int err = check_diagnostics(s, 2);
    if (err != 0) return err;
    c.advance_date();
    int day = 1 + s.current_step / nsteps;
    couple_land_atm(s, day, c.imont1, c.tmonth);
    couple_sea_atm(s, day, c);
    return 0;
}

}  // namespace orc
