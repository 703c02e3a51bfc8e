// speedy-b200: slab land / sea / sea-ice models and coupler (every step), daily forcing, and the grid-point parts
// of the model initialisation.  One thread per (grid point, member), lane = member.
//
// Reference semantics: interpolation.f90:17-94, land_model.f90:18-215, sea_model.f90:33-383, coupler.f90:13-47,
// forcing.f90:15-117, shortwave_radiation.f90:218-322, boundaries.f90:22-114, prognostics.f90:40-117.
#include "kernels.h"

namespace spdy {

namespace sf {
__device__ constexpr double AKAP = (double)(2.0f / 7.0f), CP = FL(1004.0), RGAS = AKAP * CP, GRAV = FL(9.81),
                            ALHC = FL(2501.0), SBC = FL(5.67e-8), EMISFC = FL(0.98), ALBSEA = FL(0.07),
                            ALBICE = FL(0.60), ALBSN = FL(0.60), REFRH1 = FL(0.7), DELT = 2400.0, P0 = FL(1.e+5);
__device__ constexpr double F_ASIN1 = 0x1.921fb6p+0, F_SQRT2 = 0x1.6a09e6p+0, F_LOG1013 = 0x1.a73d3ep-7;
}  // namespace sf

#define ST2D(v) (stp(c, t, c.off[v], lane) + e)

// forint / forin5 weights (interpolation.f90:39-93) evaluated per lane
struct MonthW {
    int im2, im1, i0, ip1, ip2;      // forin5 months (0-based)
    double wm2, wm1, w0, wp1, wp2;   // forin5 weights
    int j0, j2;                      // forint months (0-based)
    double wmon;
};
__device__ __forceinline__ MonthW month_weights(int imon, double tmonth) {
    MonthW w;
    int im2 = imon - 2, im1 = imon - 1, ip1 = imon + 1, ip2 = imon + 2;
    if (im2 < 1) im2 += 12;
    if (im1 < 1) im1 += 12;
    if (ip1 > 12) ip1 -= 12;
    if (ip2 > 12) ip2 -= 12;
    const double c0 = (double)(1.0f / 12.0f);
    const double t0 = c0 * tmonth, t1 = c0 * (1.0 - tmonth), t2 = 0.25 * tmonth * (1 - tmonth);
    w.wm2 = -t1 + t2, w.wm1 = -c0 + 8 * t1 - 6 * t2, w.w0 = 7 * c0 + 10 * t2, w.wp1 = -c0 + 8 * t0 - 6 * t2, w.wp2 = -t0 + t2;
    w.im2 = im2 - 1, w.im1 = im1 - 1, w.i0 = imon - 1, w.ip1 = ip1 - 1, w.ip2 = ip2 - 1;
    int imon2;
    if (tmonth <= 0.5) {
        imon2 = imon - 1;
        if (imon == 1) imon2 = 12;
        w.wmon = 0.5 - tmonth;
    } else {
        imon2 = imon + 1;
        if (imon == 12) imon2 = 1;
        w.wmon = tmonth - 0.5;
    }
    w.j0 = imon - 1, w.j2 = imon2 - 1;
    return w;
}
__device__ __forceinline__ double forin5(const double *f12, const MonthW &w, size_t lev) {
    return w.wm2 * f12[w.im2 * lev] + w.wm1 * f12[w.im1 * lev] + w.w0 * f12[w.i0 * lev] + w.wp1 * f12[w.ip1 * lev] +
           w.wp2 * f12[w.ip2 * lev];
}
__device__ __forceinline__ double forint(const double *f12, const MonthW &w, size_t lev) {
    return f12[w.j0 * lev] + w.wmon * (f12[w.j2 * lev] - f12[w.j0 * lev]);
}

// couple_land_atm + couple_sea_atm (land_model.f90:151-215, sea_model.f90:193-383).
// day0 != 0: initialisation call (coupler.f90:22-29); otherwise the per-step call (speedy.f90:72), skipped for
// members whose diagnostics check failed.
// The reference re-interpolates the monthly climatologies on every step although imont1/tmonth only change at a
// day boundary.  Here the interpolated fields (the *_obs / *_ob state arrays, which the reference stores anyway)
// are re-used while the member's date stamp SL_CPLSTAMP (set by k_control_pre) equals the current date; a host
// write to the member's state or control data clears the stamp, so the next step recomputes exactly as the
// reference would.  Identical values, 40 % less traffic on 35 of 36 steps.
__global__ void __launch_bounds__(128) k_couple(const Ctx c, const int day0) {
    using namespace sf;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (!lane_active(c, t, lane)) return;
    if (!day0 && slot(c, t, lane, SL_ERR) != 0.0) return;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    const bool full = day0 || slot(c, t, lane, SL_CPLSTAMP) != date_code(c, t, lane);
    const bool anom = slot(c, t, lane, SL_SSTACPL) != 0.0;
    const double sstfr = (double)(273.2f - 1.8f);
    double stlcl, sstcl, sicecl, ticecl, sstan_ob;
    if (full) {
        const MonthW w = month_weights((int)slot(c, t, lane, SL_IMONT1), slot(c, t, lane, SL_TMONTH));
        // ---- land climatology
        stlcl = forin5(ST2D(V_stl12), w, lev);
        const double snowdcl = forint(ST2D(V_snowd12), w, lev);
        const double soilwcl = forint(ST2D(V_soilw12), w, lev);
        *ST2D(V_stlcl_obs) = stlcl, *ST2D(V_snowdcl_obs) = snowdcl, *ST2D(V_soilwcl_obs) = soilwcl;
        *ST2D(V_snow_depth) = snowdcl, *ST2D(V_soil_avail_water) = soilwcl;
        // ---- sea climatology
        sstcl = forin5(ST2D(V_sst12), w, lev);
        sicecl = forint(ST2D(V_sea_ice_frac12), w, lev);
        const int nmon = (int)slot(c, t, lane, SL_NMONTHS);
        sstan_ob = *ST2D(V_sstan_ob);
        if (anom && nmon > 0) {  // monthly_interp (interpolation.f90:17-36)
            const double *sa = c.sst + ((size_t)c.tiles[t] * c.sst_elems) * TILE + lane + e;
            const int midx = (int)slot(c, t, lane, SL_MONTH_IDX);
            const double mf = slot(c, t, lane, SL_TMONTH);
            int imon2;
            double wmon;
            if (mf <= 0.5) imon2 = midx - 1, wmon = 0.5 - mf;
            else imon2 = midx + 1, wmon = mf - 0.5;
            const int hi = nmon + 1;
            const int a = min(max(midx, 0), hi), b = min(max(imon2, 0), hi);  // guard (the reference is unguarded)
            sstan_ob = sa[a * lev] + wmon * (sa[b * lev] - sa[a * lev]);
            *ST2D(V_sstan_ob) = sstan_ob;
        }
        if (sstcl > sstfr) {
            sicecl = fmin(0.5, sicecl);
            ticecl = sstfr;
            if (sicecl > 0.0) sstcl = sstfr + (sstcl - sstfr) / (1.0 - sicecl);
        } else {
            sicecl = fmax(0.5, sicecl);
            ticecl = sstfr + (sstcl - sstfr) / sicecl;
            sstcl = sstfr;
        }
        *ST2D(V_sstcl_ob) = sstcl, *ST2D(V_sicecl_ob) = sicecl, *ST2D(V_ticecl_ob) = ticecl;
    } else {
        stlcl = *ST2D(V_stlcl_obs);
        sstcl = *ST2D(V_sstcl_ob), sicecl = *ST2D(V_sicecl_ob), ticecl = *ST2D(V_ticecl_ob);
        sstan_ob = anom ? *ST2D(V_sstan_ob) : 0.0;
    }
    // ---- land model (land_model.f90:151-215)
    if (day0) {
        *ST2D(V_stl_lm) = stlcl, *ST2D(V_land_temp) = stlcl;
    } else if (slot(c, t, lane, SL_LANDCPL) != 0.0) {
        double tanom = *ST2D(V_stl_lm) - stlcl;
        tanom = *ST2D(V_cdland) * (tanom + *ST2D(V_rhcapl) * *ST2D(V_hfluxn));
        const double stl = tanom + stlcl;
        *ST2D(V_stl_lm) = stl, *ST2D(V_land_temp) = stl;
    } else if (full) {
        *ST2D(V_land_temp) = stlcl;
    }
    // ---- sea model
    double sst_om, tice_om, sice_om;
    if (day0) {
        sst_om = 0.0, tice_om = ticecl, sice_om = sicecl;  // sea_coupling_flag = 0 (sea_model.f90:20,262)
        *ST2D(V_wsst_ob) = 0.0;
    } else {  // run_sea_model (sea_model.f90:313-383), ice_coupling_flag = 1
        sst_om = *ST2D(V_sst_om), tice_om = *ST2D(V_tice_om);
        const double tice_am = *ST2D(V_tice_am), sice_am = *ST2D(V_sice_am);
        const double hfl2 = *(ST2D(V_hfluxn) + lev);
        const double sstfr2 = sstfr * sstfr, tam2 = tice_am * tice_am;  // x**4: products, as in physics.cu
        const double difice = (ALBSEA - ALBICE) * *ST2D(V_ssrd) + EMISFC * SBC * (sstfr2 * sstfr2 - tam2 * tam2) +
                              *(ST2D(V_shf) + lev) + *(ST2D(V_evap) + lev) * ALHC;
        const double hflux_i = hfl2 + difice * (1.0 - sice_am);
        double hflux = hfl2 - *ST2D(V_hfseacl) - sicecl * (hflux_i + FL(1.0) * (sstfr - tice_om));
        double tanom = sst_om - sstcl;
        tanom = *ST2D(V_cdsea) * (tanom + *ST2D(V_rhcaps) * hflux);
        sst_om = tanom + sstcl;
        hflux = hflux_i + FL(1.0) * (sstfr - tice_om);
        tanom = tice_om - ticecl;
        const double cdis = *ST2D(V_cdice) * (20.0 / (20.0 + fabs(tanom)));
        tanom = cdis * (tanom + *ST2D(V_rhcapi) * hflux);
        tice_om = tanom + ticecl;
        sice_om = sicecl;
    }
    *ST2D(V_sst_om) = sst_om, *ST2D(V_tice_om) = tice_om;
    const double sstan_am = anom ? sstan_ob : 0.0;
    double sst_am = sstcl + sstan_am;
    sst_am = sst_am + sice_om * (tice_om - sst_am);
    *ST2D(V_tice_am) = tice_om, *ST2D(V_sst_am) = sst_am;
    *ST2D(V_ssti_om) = sst_om + sice_om * (tice_om - sst_om);
    if (full) {  // unchanged within a day: sice_om = sicecl, sstan_am follows sstan_ob
        *ST2D(V_sice_om) = sice_om, *ST2D(V_sstan_am) = sstan_am, *ST2D(V_sice_am) = sice_om;
    }
}

// humidity.f90:44-78
__device__ __forceinline__ double qsat_raw(double ta) {
    const double e0 = 6.108e-3, c1 = FL(17.269), c2 = FL(21.875), t0 = FL(273.16), t1 = FL(35.86), t2 = FL(7.66);
    return (ta >= t0) ? e0 * exp(c1 * (ta - t0) / (ta - t1)) : e0 * exp(c2 * (ta - t0) / (ta - t2));
}

// set_forcing (forcing.f90:15-102) incl. get_zonal_average_fields / solar (shortwave_radiation.f90:218-322).
// imode 0: initialisation call for every active lane; imode 1: daily call for lanes with SL_DAILY set.
// Writes the two grid fields whose transforms are tcorh / qcorh into scratch (tcg, qcg).
__global__ void __launch_bounds__(128) k_forcing(const Ctx c, const int imode, const long long tcg, const long long qcg) {
    using namespace sf;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const size_t e = (size_t)q * TILE;
    const int j = q / IX;
    double *otc = scp(c, t, tcg, lane) + e, *oqc = scp(c, t, qcg, lane) + e;
    *otc = 0.0, *oqc = 0.0;
    if (!lane_active(c, t, lane)) return;
    if (imode == 1 && slot(c, t, lane, SL_DAILY) == 0.0) return;
    const double tyear = slot(c, t, lane, SL_TYEAR);
    if (imode == 0) {  // set_orog_land_sfc_drag (surface_fluxes.f90:324-334); radset is a shared table
        const double rhdrag = 1.0 / (GRAV * FL(2000.0));
        *ST2D(V_forog) = 1.0 + rhdrag * (1.0 - exp(-fmax(*ST2D(V_phis0), 0.0) * rhdrag));
        if (q == 0) slot(c, t, lane, SL_CO2REF) = slot(c, t, lane, SL_CO2);
    }
    // ---- get_zonal_average_fields
    {
        const double solc = FL(342.0), epssw = FL(0.020);
        const double alpha = (double)(4.0f * (float)F_ASIN1) * (tyear + (double)(10.0f / 365.0f));
        const double coz1 = 1.0 * fmax(0.0, cos(alpha - 0.0));
        const double coz2 = FL(1.8), azen = 1.0, nzen = 2.0;
        const double rzen = -cos(alpha) * FL(23.45) * F_ASIN1 / FL(90.0);
        const double fs0 = 6.0;
        // solar(tyear, 4*solc)
        const double pigr = 2.0 * F_ASIN1;
        const double al2 = 2.0 * pigr * tyear;
        const double ca1 = cos(al2), sa1 = sin(al2);
        const double ca2 = ca1 * ca1 - sa1 * sa1, sa2 = 2.0 * sa1 * ca1;
        const double ca3 = ca1 * ca2 - sa1 * sa2, sa3 = sa1 * ca2 + sa2 * ca1;
        const double decl = FL(0.006918) - FL(0.399912) * ca1 + FL(0.070257) * sa1 - FL(0.006758) * ca2 +
                            FL(0.000907) * sa2 - FL(0.002697) * ca3 + FL(0.001480) * sa3;
        const double fdis = FL(1.000110) + FL(0.034221) * ca1 + FL(0.001280) * sa1 + FL(0.000719) * ca2 + FL(0.000077) * sa2;
        const double cdecl = cos(decl), sdecl = sin(decl), tdecl = sdecl / cdecl;
        const double csolp = (4.0 * solc) / pigr;
        const double sia = c_T.sia[j], coa = c_T.coa[j];
        const double ch0 = fmin(1.0, fmax(-1.0, -tdecl * sia / coa));
        const double h0 = acos(ch0), sh0 = sin(h0);
        const double topsr = csolp * fdis * (h0 * sia * sdecl + sh0 * coa * cdecl);
        const double flat2 = FL(1.5) * (sia * sia) - 0.5;
        const double ou = 0.5 * epssw;
        const double ol = FL(0.4) * epssw * (1.0 + coz1 * sia + coz2 * flat2);
        const double zc = 1.0 + azen * pow(1.0 - (coa * cos(rzen) + sia * sin(rzen)), nzen);
        *ST2D(V_flux_solar_in) = topsr;
        *ST2D(V_zenit_correction) = zc;
        *ST2D(V_flux_ozone_upper) = topsr * ou * zc;
        *ST2D(V_flux_ozone_lower) = topsr * ol * zc;
        *ST2D(V_stratospheric_correction) = fmax(fs0 - topsr, 0.0);
    }
    // ---- surface albedo (forcing.f90:54-62)
    const double fml = *ST2D(V_fmask_land);
    {
        const double snowc = fmin(1.0, *ST2D(V_snow_depth) / FL(60.0));
        const double alb0 = *ST2D(V_alb0);
        const double alb_land = alb0 + snowc * (ALBSN - alb0);
        const double alb_sea = ALBSEA + *ST2D(V_sice_am) * (ALBICE - ALBSEA);
        *ST2D(V_snowc) = snowc, *ST2D(V_alb_land) = alb_land, *ST2D(V_alb_sea) = alb_sea;
        *ST2D(V_alb_surface) = alb_sea + fml * (alb_land - alb_sea);
    }
    if (q == 0 && slot(c, t, lane, SL_INCCO2) != 0.0)  // forcing.f90:65-72
        slot(c, t, lane, SL_CO2) = slot(c, t, lane, SL_CO2REF) * exp(FL(0.005) * (slot(c, t, lane, SL_YEAR) + tyear - 1950.0));
    // ---- orographic correction fields (forcing.f90:75-101)
    const double gamlat = 6.0 / (FL(1000.) * GRAV);
    const double corh = gamlat * *ST2D(V_phis0);
    *otc = corh;
    const double pexp = 1.0 / (RGAS * gamlat);
    const double tsfc = fml * *ST2D(V_land_temp) + *ST2D(V_fmask_sea) * *ST2D(V_sst_am);
    const double tref = tsfc + corh;
    const double psfc = pow(tsfc / tref, pexp);
    // psfc/psfc at point (1,1) of this member is the "pressure" of the sig <= 0 branch; it is 1 for finite psfc
    double qref = qsat_raw(tref);
    qref = FL(622.0) * qref / (1.0 - FL(0.378) * qref);
    double qsfc = qsat_raw(tsfc);
    qsfc = FL(622.0) * qsfc / (1.0 * psfc - FL(0.378) * qsfc);
    *oqc = REFRH1 * (qref - qsfc);
}

// masked copy of spectral fields (scratch -> state), used where a transform must not touch inactive members
__global__ void __launch_bounds__(256) k_masked_copy(const Ctx c, const FieldRef src, const FieldRef dst, const int n,
                                                     const int need_daily, const double scale) {
    const int lane = threadIdx.x & 31, t = blockIdx.y;
    if (!lane_active(c, t, lane)) return;
    if (need_daily && slot(c, t, lane, SL_DAILY) == 0.0) return;
    const double *s = refp(c, t, src, lane);
    double *d = refp(c, t, dst, lane);
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) d[(size_t)i * TILE] = s[(size_t)i * TILE] * scale;
}

// ---- initialisation: land_model_init + sea_model_init (land_model.f90:18-148, sea_model.f90:33-191) ----------
__global__ void __launch_bounds__(128) k_surface_init(const Ctx c) {
    using namespace sf;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (!lane_active(c, t, lane)) return;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    const int j = q / IX;
    const double thrsh = FL(0.1);
    const double fmo = *ST2D(V_fmask_orig);
    double fml = fmo, bml;
    if (fml >= thrsh) {
        bml = 1.0;
        if (fmo > (1.0 - thrsh)) fml = 1.0;
    } else {
        bml = 0.0, fml = 0.0;
    }
    *ST2D(V_fmask_land) = fml, *ST2D(V_bmask_land) = bml;
    double fms = 1.0 - fmo, bms;
    if (fms >= thrsh) {
        bms = 1.0;
        if (fms > (1.0 - thrsh)) fms = 1.0;
    } else {
        bms = 0.0, fms = 0.0;
    }
    *ST2D(V_fmask_sea) = fms, *ST2D(V_bmask_sea) = bms;
    // check_surface_fields (boundaries.f90:41-64): values are replaced where the binary mask is 0.
    // fill_missing_values (boundaries.f90:70-114) only acts on values < 0, absent from valid boundary files;
    // inputs with negative land/sea temperatures are rejected on the host before this kernel runs.
    const double swcap = FL(0.30), swwil = FL(0.17);
    const double veg = fmax(0.0, *ST2D(V_veg_high) + FL(0.8) * *ST2D(V_veg_low));
    const double swwil2 = 3 * swwil, rsw = 1.0 / (swcap + 3 * (swcap - swwil));
    double *stl12 = ST2D(V_stl12), *snowd12 = ST2D(V_snowd12), *soilw12 = ST2D(V_soilw12), *sst12 = ST2D(V_sst12),
           *ice12 = ST2D(V_sea_ice_frac12);
    const double *swl1 = ST2D(V_soil_wc_l1), *swl2 = ST2D(V_soil_wc_l2);
    for (int mo = 0; mo < 12; mo++) {
        if (!(bml > 0.0)) stl12[mo * lev] = 273.0, snowd12[mo * lev] = 0.0;
        const double swroot = 3 * swl2[mo * lev];
        double sw = fmin(1.0, rsw * (swl1[mo * lev] + veg * fmax(0.0, swroot - swwil2)));
        if (!(bml > 0.0)) sw = 0.0;
        soilw12[mo * lev] = sw;
        if (!(bms > 0.0)) sst12[mo * lev] = 273.0;
        double ic = fmax(ice12[mo * lev], 0.0);
        if (!(bms > 0.0)) ic = 0.0;
        ice12[mo * lev] = ic;
    }
    if ((int)slot(c, t, lane, SL_NMONTHS) > 0 && !(bms > 0.0)) {  // sea_model.f90:128: first 3 slabs only
        double *sa = c.sst + ((size_t)c.tiles[t] * c.sst_elems) * TILE + lane + e;
        for (int mo = 0; mo < 3; mo++) sa[mo * lev] = 0.0;
    }
    // heat capacities and dissipation (land_model.f90:108-145, sea_model.f90:143-188)
    const double flandmin = (double)(1.0f / 3.0f);
    const double dmask_l = (fml < flandmin) ? 0.0 : 1.0;
    const double hcapl = 1.0 * FL(2.50e+6), hcapli = 5.0 * FL(1.93e+6);
    *ST2D(V_rhcapl) = (*ST2D(V_alb0) < FL(0.4)) ? DELT / hcapl : DELT / hcapli;
    *ST2D(V_cdland) = dmask_l * 40.0 / (1.0 + dmask_l * 40.0);
    const double deglat = c_T.radang[j] * FL(90.0) / F_ASIN1;
    if ((q % IX) == 0) *(stp(c, t, c.off[V_deglat_s], lane) + (size_t)j * TILE) = deglat;
    const double crad = (double)((float)F_ASIN1 / 90.0f);
    const double coslat = cos(crad * deglat);
    const double hcaps = FL(4.18e+6) * (FL(60.) + (FL(40.) - FL(60.)) * (coslat * coslat * coslat));
    const double hcapi = FL(1.93e+6) * (FL(2.5) + (FL(1.5) - FL(2.5)) * (coslat * coslat));
    const double dmask_s = (fms < (double)(1.0f / 3.0f)) ? 0.0 : 1.0;  // l_globe: smoothed mask of ones stays one
    *ST2D(V_rhcaps) = DELT / hcaps, *ST2D(V_rhcapi) = DELT / hcapi;
    *ST2D(V_cdsea) = dmask_s * FL(90.) / (1.0 + dmask_s * FL(90.));
    *ST2D(V_cdice) = dmask_s * FL(30.0) / (1.0 + dmask_s * FL(30.0));
    *ST2D(V_hfseacl) = 0.0;
}

// initialize_boundaries / initialize_from_rest_state grid-point parts
// stage 0: phi0 = grav*orog                         (boundaries.f90:27)
// stage 1: surfg = rlog0 + rgamr*log(1 - gam2*phis0) (prognostics.f90:85-90)        -> scratch g0
// stage 2: surfg = qref*exp(qexp*surfg)              (prognostics.f90:103-107)       -> scratch g0 (in place)
__global__ void __launch_bounds__(128) k_init_grid(const Ctx c, const int stage, const long long g0) {
    using namespace sf;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const size_t e = (size_t)q * TILE;
    double *g = scp(c, t, g0, lane) + e;
    const double gam1 = 6.0 / (FL(1000.0) * GRAV);
    if (stage == 0) {
        const double v = GRAV * *ST2D(V_orog);
        if (lane_active(c, t, lane)) *ST2D(V_phi0) = v;
        *g = v;
    } else if (stage == 1) {
        const double gam2 = gam1 / 288.0, rgamr = 1.0 / (RGAS * gam1);
        *g = F_LOG1013 + rgamr * log(1.0 - gam2 * *ST2D(V_phis0));
    } else {
        const double qref = REFRH1 * FL(0.622) * 17.0, qexp = FL(7.5) / FL(2.5);
        *g = qref * exp(qexp * *g);
    }
}

// initialize_from_rest_state spectral parts (prognostics.f90:55-112).  sp_ps / sp_q: spectral scratch fields
// holding grid2spec(surfg) of stages 1 and 2.
__global__ void __launch_bounds__(128) k_init_spec(const Ctx c, const FieldRef sp_ps, const FieldRef sp_q) {
    using namespace sf;
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC || !lane_active(c, t, lane)) return;
    const int m = q % MX, n = q / MX;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE, lev = (size_t)NSP * TILE, tl = (size_t)KX * lev;
    const double trf = c.G->trfilt[q];
    const double gam1 = 6.0 / (FL(1000.0) * GRAV), rgam = RGAS * gam1, qexp = FL(7.5) / FL(2.5);
    double *vor = stp(c, t, c.off[V_vor], lane) + e, *dv = stp(c, t, c.off[V_div], lane) + e,
           *tt = stp(c, t, c.off[V_t], lane) + e, *tr = stp(c, t, c.off[V_tr], lane) + e,
           *ps = stp(c, t, c.off[V_ps], lane) + e;
    const double *phis = stp(c, t, c.off[V_phis], lane) + e;
    const double *sps = refp(c, t, sp_ps, lane) + e, *sq = refp(c, t, sp_q, lane) + e;
    (void)tl;
    for (int cc = 0; cc < 2; cc++) {
        const size_t o = cc * TILE;
        double surfs = (-gam1) * phis[o];
        if (q == 0) surfs = ((cc == 0) ? F_SQRT2 * 288.0 : 0.0 * 288.0) - gam1 * phis[o];
        for (int k = 0; k < KX; k++) {
            vor[k * lev + o] = 0.0, dv[k * lev + o] = 0.0;
            double tv = 0.0, qv = 0.0;
            if (k < 2) {
                if (q == 0) tv = (cc == 0) ? F_SQRT2 * 216.0 : 0.0 * 216.0;
            } else {
                tv = surfs * pow(c_T.fsg[k], rgam);
                qv = (sq[o] * trf) * pow(c_T.fsg[k], qexp);
            }
            tt[k * lev + o] = tv, tr[k * lev + o] = qv;
        }
        ps[o] = sps[o] * trf;
    }
}

void launch_couple(cudaStream_t s, const Ctx &c, int day0) { k_couple<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, day0); }
void launch_forcing(cudaStream_t s, const Ctx &c, int imode, long long tcg, long long qcg) {
    k_forcing<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, imode, tcg, qcg);
}
void launch_masked_copy(cudaStream_t s, const Ctx &c, FieldRef src, FieldRef dst, int n, int need_daily, double scale) {
    k_masked_copy<<<dim3(min((n + 7) / 8, 1024), c.ntiles), 256, 0, s>>>(c, src, dst, n, need_daily, scale);
}
void launch_surface_init(cudaStream_t s, const Ctx &c) { k_surface_init<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c); }
void launch_init_grid(cudaStream_t s, const Ctx &c, int stage, long long g0) {
    k_init_grid<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, stage, g0);
}
void launch_init_spec(cudaStream_t s, const Ctx &c, FieldRef sp_ps, FieldRef sp_q) {
    k_init_spec<<<dim3(NSPC / 4, c.ntiles), 128, 0, s>>>(c, sp_ps, sp_q);
}
#undef ST2D

}  // namespace spdy
