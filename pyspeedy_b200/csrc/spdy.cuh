// speedy-b200: common definitions for the CUDA hot path (sm_100a).
//
// DATA LAYOUT.  The ensemble lives in HBM as 32-member tiles ("ensemble-batched SoA"):
//   element e of variable v of member (tile, lane):  arena[(tile * ELEMS + off[v] + e) * 32 + lane]
// e is the Fortran linear index of the reference's ModelState_t array (model_state.f90, complex = re,im pair),
// so every kernel is the reference's scalar code executed by lane = member: every load/store of a warp is one
// contiguous 256-byte row, and all table look-ups are warp-uniform broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spdy_registry.h"

namespace spdy {

constexpr int TILE = 32;
constexpr int MX = 31, NX = 32, KX = 8, IX = 96, IL = 48, IY = 24, NTRUNC = 30;
constexpr int NSPC = MX * NX;       // 992 complex coefficients per spectral field
constexpr int NSP = 2 * NSPC;       // 1984 doubles per spectral field (re,im interleaved; row = 2*MX = 62)
constexpr int M2 = 2 * MX;          // 62
constexpr int NG = IX * IL;         // 4608 grid points per level
constexpr int NFOUR = M2 * IL;      // 2976 doubles per Fourier field (62 x 48)
constexpr int NSTEPS = 36, NSTRAD = 3;

// ---- per-member scalar slots (doubles; integers are stored exactly) -----------------------------------
enum Slot {
    SL_STEP = 0, SL_YEAR, SL_MONTH, SL_DAY, SL_HOUR, SL_MINUTE, SL_MONTH_IDX, SL_IMONT1, SL_TMONTH, SL_TYEAR,
    SL_CO2, SL_CO2REF, SL_INCCO2, SL_SW, SL_LANDCPL, SL_SSTACPL, SL_ERR, SL_NMONTHS, SL_INITIALIZED, SL_DAILY,
    SL_CPLSTAMP, SL_CPLDIRTY,  // coupler climatology cache (surface.cu: k_couple)
    SL_SPPT_CALLS,             // SPPT patterns generated for this member so far (noise counter; 0: first AR(1) step)
    SL_COUNT = 32
};

// ---- tables that are small and indexed with compile-time / warp-uniform indices: __constant__ -----------
struct ConstTables {
    double wa[96];   // FFTPACK twiddles (rffti1), 0-based
    double fc[4];    // taui, sqrt2, hsqt2, fft scale (double)(1.f/96.f)
    double hsg[KX + 1], dhs[KX], fsg[KX], dhsr[KX], fsgr[KX], sigl[KX], sigh[KX + 1], grdsig[KX], grdscp[KX];
    double wvi[KX][2];
    double tref[KX], tref2[KX], tref3[KX], tcorv[KX], qcorv[KX], xgeop1[KX], xgeop2[KX];
    double coriol[IL], sia[IL], coa[IL], cosgr[IL], cosgr2[IL], radang[IL], wt[IY];
    double geocorf[KX];  // lapse-rate correction factors of get_geopotential (k = 2..kx-1)
    // semi-implicit matrices for the regular leapfrog step (dt = 2*delt): constant-bank operands of the 8x8
    // mat-vecs in k_spec_step_dt (implicit.f90:234-289)
    double xc2[KX * KX], xd2[KX * KX], xj2[KX * KX * 64], dhsx2[KX];
    // column-physics constants that the reference re-derives on every call from the sigma-level tables
    // (convection.f90:63-72,84 ; vertical_diffusion.f90:62-76 ; shortwave_radiation.f90:205 ; surface_fluxes.f90:118,175)
    double ph_entrs[KX];   // entr(k) * (entmax / sum(entr))
    double ph_rdhs[KX];    // 1 / dhs(k)
    double ph_r1sig[KX];   // 1 / (1 - sigh(k+1)), k = 0..kx-2
    double ph_drh0[KX];    // rhgrad * (fsg(k) - fsg(k-1)), 0-based k >= 1
    double ph_fvdiq2[KX];  // fvdiq * sigh(k)
    double ph_sqcoa[IL];   // sqrt(coa(j))
    double ph_fm0, ph_eps1, ph_fshcq, ph_fshcse, ph_fvdiq, ph_fvdise, ph_rt1s;
};

// ---- larger tables in global memory (warp-uniform loads) ---------------------------------------------------
struct ImplTables {  // depends on the time step (implicit.f90:83-218): three instances dt/2, dt, 2dt
    double dmp1[NSPC], dmp1d[NSPC], dmp1s[NSPC], elz[NSPC];
    double xc[KX * KX], xd[KX * KX];  // (k,k1) Fortran order
    double xj[KX * KX * 64];          // (k,k1,l)
    double dhsx[KX];
};
constexpr int PQ2_KTOT = 155; // sum over m of the 4-term k-slices of even n plus those of odd n (fused_mma3.cu)
constexpr int PD2_TTOT = 93;  // sum over m of the 8-row tiles of even n plus those of odd n (fused_mma2.cu)
struct GlobTables {
    double cpol[MX * NX * IY];  // [m][n][j]  (unique half of the reference's duplicated re/im cpol)
    // parity-pure DMMA A fragments of the fused spec->grid kernel: [latitude octet 3][k-slice 155][lane 32] =
    // P(m, n, j) with n = parity + 2 * (4s + lane%4), j = 8jo + lane/4; per m the even-n slices, then the odd-n slices;
    // 0 outside the nsh2 mask
    double pq_inv2[3 * PQ2_KTOT * 32];
    // parity-pure DMMA A fragments of the fused grid->spec kernel (Gaussian weights folded in): [quad 6][tile 93][lane 32] =
    // wt(j) * P(m, n, j) with n = parity + 2 * (8 * i + lane/4), j = 4jq + lane%4; per m the even-n tiles, then the odd-n tiles
    double pq_dir2[6 * PD2_TTOT * 32];
    double el2[NSPC], elm2[NSPC], trfilt[NSPC], gradym[NSPC], gradyp[NSPC], uvdx[NSPC], uvdym[NSPC], uvdyp[NSPC],
        vddym[NSPC], vddyp[NSPC], dmp[NSPC], dmpd[NSPC], dmps[NSPC];
    double gradx[MX];
    double fband[301 * 4];
    double sppt_sigma[NSPC];  // sppt.f90:87-92: f0 * exp(-len_decorr^2 el2 / 4)
    double sppt_phi, sppt_first_fac;  // AR(1) coefficient (:30), (1 - phi^2)^(-1/2) (:95)
    ImplTables impl[3];
};

__constant__ ConstTables c_T;  // unity build: single translation unit
#define c_wa c_T.wa
#define c_fc c_T.fc

// ---- execution context passed by value to every kernel ----------------------------------------------------
struct Ctx {
    double *st;            // state arena
    double *scr;           // scratch arena (chunk-local tiles)
    double *sst;           // sst_anom arena
    const int *tiles;      // [ntiles] state tile index of chunk tile t
    const unsigned *masks; // [ntiles] active-lane mask
    const GlobTables *G;
    long long st_elems, scr_elems, sst_elems;  // doubles per lane per tile
    long long off[SPDY_NVARS];                 // element offset of each registry variable
    long long off_tcorh, off_qcorh, off_slots; // extra per-member state
    long long off_sppt;                        // SPPT AR(1) pattern, (mx,nx,kx) complex
    int ntiles;
    int sst_months;                            // slabs per member in the sst arena
    const int *outer_zero;                     // [2 * ntiles] or nullptr.  Non-zero: the time filter is the identity on every
                                               // coefficient outside the triangular truncation (vor, div, t, tr, ps) of the
                                               // tile's active members (k_scan_outer at the start of a multi-step call)
    int diag_out;                              // 0: an intermediate step of a multi-step driver call -- the column physics
                                               // does not store the outputs that the NEXT step overwrites before anything on
                                               // the device (or the host: the call has not returned) can read them
};

__device__ __forceinline__ double *stp(const Ctx &c, int t, long long off, int lane) {
    return c.st + ((long long)c.tiles[t] * c.st_elems + off) * TILE + lane;
}
// The tile list is never written while a kernel runs: a non-coherent (invariant) load lets the compiler keep the tile index
// in a register across stores instead of re-reading it for every state address (42 reloads in k_physics before this: 0.543
// -> 0.516 ms per intermediate step at 512 members).  The short bandwidth-bound kernels keep stp: there the hoisted loads
// cost registers (k_couple 48 -> 64, k_spec_step_vq 32 -> 42) and 3-5 % of their time.
__device__ __forceinline__ double *stp_nc(const Ctx &c, int t, long long off, int lane) {
    return c.st + ((long long)__ldg(c.tiles + t) * c.st_elems + off) * TILE + lane;
}
__device__ __forceinline__ double *scp(const Ctx &c, int t, long long off, int lane) {
    return c.scr + ((long long)t * c.scr_elems + off) * TILE + lane;
}
__device__ __forceinline__ bool lane_active(const Ctx &c, int t, int lane) { return (c.masks[t] >> lane) & 1u; }
__device__ __forceinline__ double &slot(const Ctx &c, int t, int lane, int s) {
    return *stp(c, t, c.off_slots + s, lane);
}

// A field reference: element offset inside a tile; bit 62 selects the scratch arena.
typedef long long FieldRef;
constexpr long long REF_SCR = 1ll << 62;
__device__ __forceinline__ double *refp(const Ctx &c, int t, FieldRef r, int lane) {
    return (r & REF_SCR) ? scp(c, t, r & ~REF_SCR, lane) : stp(c, t, r, lane);
}

// Ampere-style asynchronous 16-byte global -> shared copies (L1 bypass)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void prefetch_l2(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// default-REAL literal widened to double, e.g. FL(0.05) == (double)0.05f
#define FL(x) ((double)(x##f))

// ---- scratch arena layout (element offsets; per lane) ------------------------------------------------------
struct ScratchLayout {
    // grid-point fields (NG * KX each unless noted)
    long long ug, vg, tg, vorg, divg, trg;           // dynamics, time level j2
    long long ptg, pqg, pphig, pug8, pvg8, pslg;    // physics, time level j1 (u,v: lowest level only)
    long long px, py, psdtg;                         // 2-D
    long long utend, vtend, ttend, trtend;
    // spectral fields
    long long ucos, vcos, ucosp8, vcosp8, dpx, dpy;  // inverse-transform inputs
    long long sfwd;                                  // 73 forward-transform outputs
    // Fourier fields (62 x 48)
    long long four;                                  // max(77, 73) fields
    long long diagp;                                 // partial sums of the diagnostics check
    long long total;                                 // doubles per lane in use (total_base, or total_sppt when SPPT is on)
    // SPPT (allocated only while it is switched on): pattern on the grid, copies of the dynamical tendencies
    long long spptg, tdyn, qdyn, udyn8, vdyn8, total_base, total_sppt;
};
ScratchLayout make_scratch_layout();

// forward-transform output slots inside `sfwd` (each NSP doubles): per level k = 0..7
constexpr int FW_SU = 0, FW_SV = 8, FW_KE = 16, FW_UT = 24, FW_VT = 32, FW_TT = 40, FW_UQ = 48, FW_VQ = 56,
              FW_QT = 64, FW_PS = 72, FW_COUNT = 73;

}  // namespace spdy
