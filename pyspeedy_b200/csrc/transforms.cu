// speedy-b200: spectral transform kernels (Legendre + 96-point real FFT) and spectral-space operators.
//
// Reference semantics: legendre.f90:130-221, fourier.f90:63-123 (+ fftpack.f90), spectral.f90:134-296.
// Layout: lane = ensemble member (32 per tile), see spdy.cuh.  All kernels: gridDim.y = chunk tiles.
//
//   k_legendre_inv : spectral (62x32) -> Fourier (62x48).  Per zonal wavenumber m a (24 x n) contraction; the
//                    Legendre polynomial P(m,n,j) is warp-uniform (one 128-bit broadcast load feeds 2 j's for
//                    32 members x {re,im}); register tile = 6 latitudes x {even,odd} x {re,im} = 24 accumulators.
//   k_fft_inv      : Fourier -> grid, generated butterfly items (fft96_gen.cuh), 2 line-groups per CTA,
//                    one shared-memory exchange between the two register stages.
//   k_fft_fwd<M>   : grid -> Fourier with the grid-point products fused into the loads (M = loader mode).
//   k_legendre_dir : Fourier -> spectral, Gaussian quadrature (fold N/S, weight, 24-term dot products).
#include "kernels.h"

namespace spdy {


#define FFT_LS TILE
#ifndef FFT_MINBLOCKS
#define FFT_MINBLOCKS 3
#endif
#include "fft96_gen.cuh"

// ------------------------------------------------------------------------------------------- Legendre inverse
// One warp = one (field, m-pair); pairs (m, 30-m) balance the triangular truncation: 34 n-terms per pair.
// What bounds this kernel is the L1/LSU data pipe, not FP64 or HBM (tools/microbench/l1pipe.cu: a warp-uniform
// value costs 2 SM-cycles through LDG but 1 through LDS; a 256-byte member row costs 2 either way).  The warp
// therefore copies the Legendre polynomials of its pair (34 rows x 24 latitudes = 6.5 KB, contiguous in cpol)
// into a private shared-memory slab with cp.async and feeds the FMAs from broadcast LDS.128; the spectral
// coefficients stream from global memory (read-only path), 6 latitudes x {even,odd} x {re,im} register tile.
constexpr int LEGI_ROWS = 34;
constexpr int LEGI_SMEM = 4 * LEGI_ROWS * IY * 8;  // 26,112 bytes per CTA (4 warps)
__device__ __forceinline__ void leg_stage_rows(double *__restrict__ S, const double *__restrict__ Pm, int nrows, int lane) {
    const int pieces = nrows * (IY / 2);  // 16-byte pieces, rows are contiguous in cpol[m][n][j]
    for (int i = lane; i < pieces; i += 32) cp_async16(S + 2 * i, Pm + 2 * i);
}
__device__ __forceinline__ void leg_inv_one_m(const double *__restrict__ X, double *__restrict__ F,
                                              const double *__restrict__ Ps, const int m0) {
    const int nmax = 31 - m0;  // n0 = 0..nmax are inside the nsh2 mask (legendre.f90:68-77)
    const double *Xr = X + (2 * m0) * TILE, *Xi = Xr + TILE;
#pragma unroll 1
    for (int jt = 0; jt < 4; jt++) {
        double er[6], ei[6], orr[6], oi[6];
#pragma unroll
        for (int q = 0; q < 6; q++) er[q] = ei[q] = orr[q] = oi[q] = 0.0;
#pragma unroll 4
        for (int n0 = 0; n0 <= nmax; n0 += 2) {  // even parity: l - m even
            const double xr = __ldg(Xr + (size_t)n0 * M2 * TILE), xi = __ldg(Xi + (size_t)n0 * M2 * TILE);
            const double2 *p = reinterpret_cast<const double2 *>(Ps + n0 * IY + jt * 6);
            const double2 p0 = p[0], p1 = p[1], p2 = p[2];
            er[0] += xr * p0.x, ei[0] += xi * p0.x, er[1] += xr * p0.y, ei[1] += xi * p0.y;
            er[2] += xr * p1.x, ei[2] += xi * p1.x, er[3] += xr * p1.y, ei[3] += xi * p1.y;
            er[4] += xr * p2.x, ei[4] += xi * p2.x, er[5] += xr * p2.y, ei[5] += xi * p2.y;
        }
#pragma unroll 4
        for (int n0 = 1; n0 <= nmax; n0 += 2) {  // odd parity
            const double xr = __ldg(Xr + (size_t)n0 * M2 * TILE), xi = __ldg(Xi + (size_t)n0 * M2 * TILE);
            const double2 *p = reinterpret_cast<const double2 *>(Ps + n0 * IY + jt * 6);
            const double2 p0 = p[0], p1 = p[1], p2 = p[2];
            orr[0] += xr * p0.x, oi[0] += xi * p0.x, orr[1] += xr * p0.y, oi[1] += xi * p0.y;
            orr[2] += xr * p1.x, oi[2] += xi * p1.x, orr[3] += xr * p1.y, oi[3] += xi * p1.y;
            orr[4] += xr * p2.x, oi[4] += xi * p2.x, orr[5] += xr * p2.y, oi[5] += xi * p2.y;
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const int j0 = jt * 6 + q, jn = IL - 1 - j0;  // legendre.f90:163-167: row il+1-j gets even+odd
            double *fn = F + ((size_t)jn * M2 + 2 * m0) * TILE, *fs = F + ((size_t)j0 * M2 + 2 * m0) * TILE;
            fn[0] = er[q] + orr[q], fn[TILE] = ei[q] + oi[q];
            fs[0] = er[q] - orr[q], fs[TILE] = ei[q] - oi[q];
        }
    }
}

__global__ void __launch_bounds__(128) k_legendre_inv(const Ctx c, const InvDesc *__restrict__ descs, long long four_off) {
    extern __shared__ __align__(16) double leg_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x >> 2, unit = (blockIdx.x & 3) * 4 + warp;  // 16 units per field
    const int t = blockIdx.y;
    const double *X = refp(c, t, descs[f].src, lane);
    double *F = scp(c, t, four_off + (long long)f * NFOUR, lane);
    const double *P = c.G->cpol;
    double *S = leg_sm + warp * (LEGI_ROWS * IY);
    const int ma = (unit < 15) ? unit : 15, mb = 30 - unit;
    double *Sb = S + (32 - ma) * IY;
    leg_stage_rows(S, P + (size_t)ma * NX * IY, 32 - ma, lane);
    if (unit < 15) leg_stage_rows(Sb, P + (size_t)mb * NX * IY, 32 - mb, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    leg_inv_one_m(X, F, S, ma);
    if (unit < 15) leg_inv_one_m(X, F, Sb, mb);
}

// -------------------------------------------------------------------------------------------- Legendre direct
// CTA = the wavenumber pair (m, 30-m) of one field; the two warps of a wavenumber take one parity each (n - m even
// uses only the N+S fold, odd only the N-S fold), so every Legendre value fetched feeds {re,im} of 32 members.
// The pair's polynomials are staged in shared memory as above; the 48 Fourier rows of the wavenumber are loaded
// straight into registers (independent loads, all in flight), folded and weighted (legendre.f90:196-197), then
// 24-term dot products run two spectral rows (four accumulation chains) at a time (legendre.f90:206-218; rows
// beyond the nsh2 mask are written as zero).
constexpr int LEGD_SMEM = LEGI_ROWS * IY * 8;  // 6,528 bytes per CTA
__global__ void __launch_bounds__(128) k_legendre_dir(const Ctx c, const FwdOut *__restrict__ outs, long long four_off) {
    __shared__ __align__(16) double Ps[LEGI_ROWS * IY];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x >> 4, unit = blockIdx.x & 15, sel = warp >> 1, par = warp & 1;
    const int t = blockIdx.y;
    const bool valid = !(unit == 15 && sel == 1);
    const int ma = (unit < 15) ? unit : 15, mb = 30 - unit;
    const int m0 = sel ? mb : ma;
    const double *P = c.G->cpol;
    // warps 0,1 stage the rows of ma, warps 2,3 those of mb (half of the 16-byte pieces each)
    if (valid) {
        const int nrows = 32 - m0, pieces = nrows * (IY / 2);
        double *S = Ps + (sel ? (32 - ma) * IY : 0);
        const double *g = P + (size_t)m0 * NX * IY;
        for (int i = par * 32 + lane; i < pieces; i += 64) cp_async16(S + 2 * i, g + 2 * i);
    }
    cp_async_commit();
    double wr[IY], wi[IY];
    if (valid) {
        const double *Fm = scp(c, t, four_off + (long long)f * NFOUR, lane) + (size_t)(2 * m0) * TILE;
#pragma unroll
        for (int j = 0; j < IY; j++) {
            const double *ps = Fm + (size_t)j * M2 * TILE, *pn = Fm + (size_t)(IL - 1 - j) * M2 * TILE;
            const double sr = par ? -ps[0] : ps[0], si = par ? -ps[TILE] : ps[TILE];
            wr[j] = (pn[0] + sr) * c_T.wt[j];
            wi[j] = (pn[TILE] + si) * c_T.wt[j];
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (!valid) return;
    double *X = refp(c, t, outs[f].dst, lane) + (size_t)(2 * m0) * TILE;
    const double *Pm = Ps + (sel ? (32 - ma) * IY : 0);
    const int nmax = min(30, 31 - m0);
#pragma unroll 1
    for (int n0 = par; n0 < NX; n0 += 4) {
        double ar0 = 0.0, ai0 = 0.0, ar1 = 0.0, ai1 = 0.0;
        if (n0 <= nmax) {
            const bool two = (n0 + 2 <= nmax);  // row n0+2 may lie outside the staged slab: re-use row n0, discard
            const double2 *p0 = reinterpret_cast<const double2 *>(Pm + n0 * IY), *p1 = two ? p0 + IY : p0;
#pragma unroll
            for (int j = 0; j < IY; j += 2) {
                const double2 a = p0[j >> 1], b = p1[j >> 1];
                ar0 += a.x * wr[j], ai0 += a.x * wi[j], ar1 += b.x * wr[j], ai1 += b.x * wi[j];
                ar0 += a.y * wr[j + 1], ai0 += a.y * wi[j + 1], ar1 += b.y * wr[j + 1], ai1 += b.y * wi[j + 1];
            }
            if (!two) ar1 = 0.0, ai1 = 0.0;
        }
        double *x0 = X + (size_t)n0 * M2 * TILE, *x1 = x0 + (size_t)2 * M2 * TILE;
        x0[0] = ar0, x0[TILE] = ai0, x1[0] = ar1, x1[TILE] = ai1;
    }
}

// ----------------------------------------------------------------------------------------------- inverse FFT
// CTA = 4 warps x 2 line-groups (a line-group = 32 members x one latitude of one field).  Every warp has a STATIC
// list of stage-A items (balanced by flop count: A0 24, A1..A5 96, A6 36) written out as straight-line code, so all
// of its global loads (read-only path) are independent of each other and of the shared-memory stores and can be in
// flight together; one barrier; then a static list of stage-B items.
struct LdFour {
    const double *p;
    __device__ __forceinline__ double operator()(int r) const { return __ldg(p + r * TILE); }
};
struct StGrid {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int i, double v) const { p[i * TILE] = v * sc; }
};

__global__ void __launch_bounds__(128, FFT_MINBLOCKS) k_fft_inv(const Ctx c, const InvDesc *__restrict__ descs,
                                                                 long long four_off, int nlg) {
    __shared__ double sm[2 * IX * TILE];  // 48 KB: exchange buffers of 2 line-groups
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.y, lg0 = blockIdx.x * 2;  // nlg is even (48 latitudes per field)
    const int f0 = lg0 / IL, j0 = lg0 - f0 * IL, f1 = (lg0 + 1) / IL, j1 = (lg0 + 1) - f1 * IL;
    const LdFour ld0{scp(c, t, four_off + (long long)f0 * NFOUR + j0 * M2, lane)};
    const LdFour ld1{scp(c, t, four_off + (long long)f1 * NFOUR + j1 * M2, lane)};
    double *s0 = sm + lane, *s1 = sm + IX * TILE + lane;
    if (warp == 0) {
        fftb_A1(ld0, s0), fftb_A2(ld0, s0), fftb_A3(ld0, s0);
    } else if (warp == 1) {
        fftb_A4(ld0, s0), fftb_A5(ld0, s0), fftb_A1(ld1, s1);
    } else if (warp == 2) {
        fftb_A2(ld1, s1), fftb_A3(ld1, s1), fftb_A0(ld0, s0), fftb_A6(ld0, s0);
    } else {
        fftb_A4(ld1, s1), fftb_A5(ld1, s1), fftb_A0(ld1, s1), fftb_A6(ld1, s1);
    }
    __syncthreads();
    const InvDesc d0 = descs[f0], d1 = descs[f1];
    const StGrid st0{scp(c, t, d0.dst + (long long)j0 * IX, lane), d0.kcos == 1 ? 1.0 : c_T.cosgr[j0]};
    const StGrid st1{scp(c, t, d1.dst + (long long)j1 * IX, lane), d1.kcos == 1 ? 1.0 : c_T.cosgr[j1]};
    if (warp == 0) {
        fftb_B0(s0, st0), fftb_B1(s0, st0), fftb_B2(s0, st0), fftb_B3(s0, st0);
    } else if (warp == 1) {
        fftb_B4(s0, st0), fftb_B5(s0, st0), fftb_B6(s0, st0), fftb_B7(s0, st0);
    } else if (warp == 2) {
        fftb_B0(s1, st1), fftb_B1(s1, st1), fftb_B2(s1, st1), fftb_B3(s1, st1);
    } else {
        fftb_B4(s1, st1), fftb_B5(s1, st1), fftb_B6(s1, st1), fftb_B7(s1, st1);
    }
}

// ----------------------------------------------------------------------------------------------- forward FFT
// Loader modes: the grid-point products of tendencies.f90:238-268 and the cos(lat) scalings of
// spectral.f90:229-242 are applied while loading, so these fields never exist in HBM.
template <int MODE> struct LdGrid {
    const double *a, *b;
    double k0, sc;
    __device__ __forceinline__ double operator()(int i) const {
        if (MODE == FM_PLAIN) return __ldg(a + i * TILE);
        if (MODE == FM_COS) return __ldg(a + i * TILE) * sc;
        if (MODE == FM_KE) {
            const double u = __ldg(a + i * TILE), v = __ldg(b + i * TILE);
            return 0.5 * (u * u + v * v);
        }
        if (MODE == FM_FLUXT) return (-__ldg(a + i * TILE) * (__ldg(b + i * TILE) - k0)) * sc;
        return (-__ldg(a + i * TILE) * __ldg(b + i * TILE)) * sc;  // FM_FLUX
    }
};
struct StFour {
    double *p;
    double scale;
    __device__ __forceinline__ void operator()(int r, double v) const { p[r * TILE] = v * scale; }
};

template <int MODE>
__device__ __forceinline__ LdGrid<MODE> make_ld(const Ctx &c, int t, const FwdDesc &d, int j, int lane) {
    LdGrid<MODE> ld;
    ld.a = refp(c, t, d.a, lane) + (size_t)j * IX * TILE;
    ld.b = (MODE == FM_KE || MODE == FM_FLUXT || MODE == FM_FLUX) ? refp(c, t, d.b, lane) + (size_t)j * IX * TILE : nullptr;
    ld.k0 = d.k0;
    ld.sc = (d.kcos == 3) ? c_T.cosgr2[j] : c_T.cosgr[j];
    return ld;
}

// CTA = 2 line-groups x 2 warps.  Stage-A item kk is item 0 on inputs kk + 8q with outputs 12kk.. (the eight
// 12-point problems of passes 1+2 are the same DAG), so each warp runs ONE inlined body four times with shifted
// pointers; stage B has seven different items, split 4 + 3 between the two warps of a line-group.  Compared with a
// fully static schedule (30 inlined bodies, 52 KB of code) this removes the instruction-cache stalls that ncu
// showed as 20 % of the samples of this kernel.
#ifndef FFTF_UNROLL
#define FFTF_UNROLL 2
#endif
constexpr int kFftfUnroll = FFTF_UNROLL;
template <int MODE>
__global__ void __launch_bounds__(128, FFT_MINBLOCKS) k_fft_fwd(const Ctx c, const FwdDesc *__restrict__ descs,
                                                                 long long four_off, int nlg) {
    __shared__ double sm[2 * IX * TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lgi = warp >> 1, half = warp & 1;
    const int t = blockIdx.y, lg = blockIdx.x * 2 + lgi;  // nlg is even
    const int f = lg / IL, j = lg - f * IL;
    const FwdDesc d = descs[f];
    double *s = sm + lgi * (IX * TILE) + lane;
    {
        const LdGrid<MODE> ld = make_ld<MODE>(c, t, d, j, lane);
#pragma unroll kFftfUnroll
        for (int kk = half * 4; kk < half * 4 + 4; kk++) {
            LdGrid<MODE> l2 = ld;
            l2.a += kk * TILE;
            if (MODE == FM_KE || MODE == FM_FLUXT || MODE == FM_FLUX) l2.b += kk * TILE;
            fftf_A0(l2, s + 12 * kk * TILE);
        }
    }
    __syncthreads();
    // stage B: B0 21, B1..B5 102, B6 34 flops; item B0 owns row 0 and zeroes row 1 (Im of m = 0, fourier.f90:117)
    double *four = scp(c, t, four_off + (long long)d.fidx * NFOUR + j * M2, lane);
    const StFour st{four, c_T.fc[3]};
    if (half == 0) {
        fftf_B1(s, st), fftf_B2(s, st), fftf_B3(s, st), fftf_B0(s, st);
        four[TILE] = 0.0;
    } else {
        fftf_B4(s, st), fftf_B5(s, st), fftf_B6(s, st);
    }
}

// ------------------------------------------------------------------------- whole-line transforms in registers
// One thread transforms one (line, member): the 96-value exchange between the two register stages stays in the
// thread's registers (fft96_reg_gen.cuh), no shared memory, no barrier.  ~210-230 registers, no spills.
template <class LD, class ST> __device__ __forceinline__ void fftb_line(const LD ld, const ST st) {
    double x[IX];
    rfftb_A0(ld, x), rfftb_A1(ld, x), rfftb_A2(ld, x), rfftb_A3(ld, x), rfftb_A4(ld, x), rfftb_A5(ld, x), rfftb_A6(ld, x);
    rfftb_B0(x, st), rfftb_B1(x, st), rfftb_B2(x, st), rfftb_B3(x, st), rfftb_B4(x, st), rfftb_B5(x, st), rfftb_B6(x, st),
        rfftb_B7(x, st);
}
template <class LD, class ST> __device__ __forceinline__ void fftf_line(const LD ld, const ST st) {
    double x[IX];
    rfftf_A0(ld, x), rfftf_A1(ld, x), rfftf_A2(ld, x), rfftf_A3(ld, x), rfftf_A4(ld, x), rfftf_A5(ld, x), rfftf_A6(ld, x),
        rfftf_A7(ld, x);
    rfftf_B0(x, st), rfftf_B1(x, st), rfftf_B2(x, st), rfftf_B3(x, st), rfftf_B4(x, st), rfftf_B5(x, st), rfftf_B6(x, st);
}
#ifdef FFT_REG
__global__ void __launch_bounds__(128, 2) k_fft_inv_reg(const Ctx c, const InvDesc *__restrict__ descs, long long four_off, int nlg) {
    const int lane = threadIdx.x & 31, t = blockIdx.y, lg = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (lg >= nlg) return;
    const int f = lg / IL, j = lg - f * IL;
    const LdFour ld{scp(c, t, four_off + (long long)f * NFOUR + j * M2, lane)};
    const InvDesc d = descs[f];
    const StGrid st{scp(c, t, d.dst + (long long)j * IX, lane), d.kcos == 1 ? 1.0 : c_T.cosgr[j]};
    fftb_line(ld, st);
}
template <int MODE>
__global__ void __launch_bounds__(128, 2) k_fft_fwd_reg(const Ctx c, const FwdDesc *__restrict__ descs, long long four_off, int nlg) {
    const int lane = threadIdx.x & 31, t = blockIdx.y, lg = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (lg >= nlg) return;
    const int f = lg / IL, j = lg - f * IL;
    const FwdDesc d = descs[f];
    const LdGrid<MODE> ld = make_ld<MODE>(c, t, d, j, lane);
    double *four = scp(c, t, four_off + (long long)d.fidx * NFOUR + j * M2, lane);
    fftf_line(ld, StFour{four, c_T.fc[3]});
    four[TILE] = 0.0;  // fourier.f90:117
}
#endif

// ------------------------------------------------------------------------------ spectral-space pre-operators
// uvspec (spectral.f90:190-214) for nlev levels, one warp per complex coefficient (m,n), lane = member
__device__ __forceinline__ void uvspec_elem(const GlobTables *G, const double *vor, const double *dv, double *u,
                                            double *v, int m, int n) {
    const int q = m + MX * n;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE, up = (size_t)M2 * TILE;
    // zp = uvdx*vor*(0,1) ; zc = uvdx*div*(0,1)
    const double ux = G->uvdx[q];
    const double zpr = -(ux * vor[e + TILE]), zpi = ux * vor[e];
    const double zcr = -(ux * dv[e + TILE]), zci = ux * dv[e];
    if (n == 0) {
        const double yp = G->uvdyp[q];
        u[e] = zcr - yp * vor[e + up], u[e + TILE] = zci - yp * vor[e + up + TILE];
        v[e] = zpr + yp * dv[e + up], v[e + TILE] = zpi + yp * dv[e + up + TILE];
    } else if (n == NX - 1) {
        const double ym = G->uvdym[q];
        u[e] = ym * vor[e - up], u[e + TILE] = ym * vor[e - up + TILE];
        v[e] = -ym * dv[e - up], v[e + TILE] = -ym * dv[e - up + TILE];
    } else {
        const double ym = G->uvdym[q], yp = G->uvdyp[q];
        v[e] = (-ym * dv[e - up] + yp * dv[e + up]) + zpr;
        v[e + TILE] = (-ym * dv[e - up + TILE] + yp * dv[e + up + TILE]) + zpi;
        u[e] = (ym * vor[e - up] - yp * vor[e + up]) + zcr;
        u[e + TILE] = (ym * vor[e - up + TILE] - yp * vor[e + up + TILE]) + zci;
    }
}

// tri: the consumer is an inverse transform that reads the rows n <= 31 - m only (legendre.f90:143-158 through nsh2;
// k_spec2grid_mma4 never touches the others), so the 465 coefficients beyond them are neither loaded nor stored
__global__ void __launch_bounds__(128) k_uvspec(const Ctx c, FieldRef vor, FieldRef dv, FieldRef u, FieldRef v, int nlev,
                                                int tri) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    if (tri && m + n > NTRUNC + 1) return;
    const double *pv = refp(c, t, vor, lane), *pd = refp(c, t, dv, lane);
    double *pu = refp(c, t, u, lane), *pw = refp(c, t, v, lane);
    for (int k = 0; k < nlev; k++)
        uvspec_elem(c.G, pv + (size_t)k * NSP * TILE, pd + (size_t)k * NSP * TILE, pu + (size_t)k * NSP * TILE,
                    pw + (size_t)k * NSP * TILE, m, n);
}

// gradient (spectral.f90:275-296)
// gradient (spectral.f90:275-296) at coefficient (m,n)
__device__ __forceinline__ void gradient_elem(const GlobTables *G, const double *p, double *px, double *py, int m, int n) {
    const int q = m + MX * n;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE, up = (size_t)M2 * TILE;
    const double gx = G->gradx[m];
    px[e] = -(gx * p[e + TILE]);
    px[e + TILE] = gx * p[e];
    if (n == 0) {
        py[e] = G->gradyp[q] * p[e + up], py[e + TILE] = G->gradyp[q] * p[e + up + TILE];
    } else if (n == NX - 1) {
        py[e] = -G->gradym[q] * p[e - up], py[e + TILE] = -G->gradym[q] * p[e - up + TILE];
    } else {
        py[e] = -G->gradym[q] * p[e - up] + G->gradyp[q] * p[e + up];
        py[e + TILE] = -G->gradym[q] * p[e - up + TILE] + G->gradyp[q] * p[e + up + TILE];
    }
}
__global__ void __launch_bounds__(128) k_gradient(const Ctx c, FieldRef psi, FieldRef dx, FieldRef dy, int tri) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    if (tri && m + n > NTRUNC + 1) return;  // see k_uvspec
    gradient_elem(c.G, refp(c, t, psi, lane), refp(c, t, dx, lane), refp(c, t, dy, lane), m, n);
}

// geopotential (geopotential.f90:36-77): phi from T(time level) and phis ; one warp per coefficient
// set_geopotential (geopotential.f90:36-77) at coefficient (m,n): hydrostatic integration + lapse-rate correction on m = 0
__device__ __forceinline__ void geopotential_elem(const double *T, const double *ps, double *ph, const int m, const bool act) {
    const size_t lev = (size_t)NSP * TILE;
#pragma unroll
    for (int cc = 0; cc < 2; cc++) {
        double tk[KX], p[KX];
#pragma unroll
        for (int k = 0; k < KX; k++) tk[k] = T[k * lev + cc * TILE];
        p[KX - 1] = ps[cc * TILE] + c_T.xgeop1[KX - 1] * tk[KX - 1];
#pragma unroll
        for (int k = KX - 2; k >= 0; k--) p[k] = (p[k + 1] + c_T.xgeop2[k + 1] * tk[k + 1]) + c_T.xgeop1[k] * tk[k];
        if (m == 0) {
#pragma unroll
            for (int k = 1; k < KX - 1; k++) p[k] = p[k] + c_T.geocorf[k] * (tk[k + 1] - tk[k - 1]);
        }
        if (act) {
#pragma unroll
            for (int k = 0; k < KX; k++) ph[k * lev + cc * TILE] = p[k];
        }
    }
}
__global__ void __launch_bounds__(128) k_geopotential(const Ctx c, FieldRef tref_, FieldRef phis, FieldRef phi) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC) return;
    const int m = q % MX, n = q / MX;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE;
    geopotential_elem(refp(c, t, tref_, lane) + e, refp(c, t, phis, lane) + e, refp(c, t, phi, lane) + e, m, lane_active(c, t, lane));
}

// ------------------------------------------------------------------------------------------------- launchers
void launch_legendre_inv(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf, long long four_off) {
    if (!nf) return;
    k_legendre_inv<<<dim3(nf * 4, c.ntiles), 128, LEGI_SMEM, s>>>(c, d, four_off);
}
void launch_fft_inv(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf, long long four_off) {
    const int nlg = nf * IL;
#ifdef FFT_REG
    if (nf) k_fft_inv_reg<<<dim3((nlg + 3) / 4, c.ntiles), 128, 0, s>>>(c, d, four_off, nlg);
#else
    if (nf) k_fft_inv<<<dim3((nlg + 1) / 2, c.ntiles), 128, 0, s>>>(c, d, four_off, nlg);
#endif
}
void launch_fft_fwd(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, int nf, long long four_off) {
    if (!nf) return;
    const int nlg = nf * IL;
#ifdef FFT_REG
    const dim3 g((nlg + 3) / 4, c.ntiles);
    switch (mode) {
        case FM_PLAIN: k_fft_fwd_reg<FM_PLAIN><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_COS: k_fft_fwd_reg<FM_COS><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_KE: k_fft_fwd_reg<FM_KE><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_FLUXT: k_fft_fwd_reg<FM_FLUXT><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        default: k_fft_fwd_reg<FM_FLUX><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
    }
}
#else
    const dim3 g((nlg + 1) / 2, c.ntiles);
    switch (mode) {
        case FM_PLAIN: k_fft_fwd<FM_PLAIN><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_COS: k_fft_fwd<FM_COS><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_KE: k_fft_fwd<FM_KE><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        case FM_FLUXT: k_fft_fwd<FM_FLUXT><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
        default: k_fft_fwd<FM_FLUX><<<g, 128, 0, s>>>(c, d, four_off, nlg); break;
    }
}
#endif
void launch_legendre_dir(cudaStream_t s, const Ctx &c, const FwdOut *o, int nf, long long four_off) {
    if (!nf) return;
    k_legendre_dir<<<dim3(nf * 16, c.ntiles), 128, 0, s>>>(c, o, four_off);
}
void launch_uvspec(cudaStream_t s, const Ctx &c, FieldRef vor, FieldRef dv, FieldRef u, FieldRef v, int nlev, int tri) {
    k_uvspec<<<dim3(NSPC / 4, c.ntiles), 128, 0, s>>>(c, vor, dv, u, v, nlev, tri);
}
void launch_gradient(cudaStream_t s, const Ctx &c, FieldRef psi, FieldRef dx, FieldRef dy, int tri) {
    k_gradient<<<dim3(NSPC / 4, c.ntiles), 128, 0, s>>>(c, psi, dx, dy, tri);
}
void launch_geopotential(cudaStream_t s, const Ctx &c, FieldRef tlev, FieldRef phis, FieldRef phi) {
    k_geopotential<<<dim3(NSPC / 4, c.ntiles), 128, 0, s>>>(c, tlev, phis, phi);
}
void upload_const_tables(const ConstTables &C) { cudaMemcpyToSymbol(c_T, &C, sizeof(C)); }

}  // namespace spdy
