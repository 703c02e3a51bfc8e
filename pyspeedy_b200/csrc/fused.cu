// speedy-b200: FUSED spectral transforms -- Legendre and 96-point FFT in one kernel, the Fourier array (62 x 48)
// never leaves the SM (SURVEY.md 8d rung (ii)): spec2grid moves 52.7 KB per field through HBM instead of 100.4 KB.
//
// Reference semantics are those of transforms.cu (legendre.f90:130-221, fourier.f90:63-123, fftpack.f90); the same
// generated butterfly items (fft96_gen.cuh) are used.
//
// Work item = (field, tile, half): the 16 members h*16..h*16+15 of a tile (global rows are still the 256-byte member
// rows, a half-tile touches 128-byte segments).  CTA = 12 warps, persistent over work items:
//   * 8 "L" warps (Legendre): lane = (component re/im, member).  Each warp owns 4 zonal wavenumbers (two balanced
//     pairs (m, 30-m)) and keeps THEIR spectral coefficients (inverse) or accumulators (direct) in registers for all 24
//     latitude pairs: 68 doubles per lane.  The Legendre polynomials of the current latitude pair come from a
//     double-buffered 7.9 KB shared-memory slice (table layout [m][j][n], 128-bit broadcast loads).
//   * 4 "F" warps (FFT): lane = (hemisphere, member), i.e. a warp transforms the 32 lines of one latitude pair.
//     Stage A -> 24.6 KB exchange buffer -> stage B.
//   * L and F warps form a producer/consumer pipeline over latitude pairs through a double-buffered Fourier buffer
//     (2 x 15.9 KB) with named barriers (FULL/EMPTY), so the Legendre FMAs, the butterflies and the HBM traffic of
//     consecutive latitude pairs overlap.  The forward kernel stages its grid rows with cp.async one pair ahead.
#include "kernels.h"

namespace spdy {

constexpr int FT_L_WARPS = 8, FT_NL = 32 * FT_L_WARPS;
constexpr int FS_ELEMS = M2 * TILE;      // Fourier buffer of one latitude pair: 62 rows x (2 hemispheres x 16 members)
constexpr int ES_ELEMS = IX * TILE;      // FFT exchange buffer (private to one F warp)
constexpr int PS_ELEMS = MX * NX;        // Legendre slice of one latitude pair
constexpr int GS_ELEMS = IX * TILE;      // staged grid rows of one latitude pair (forward)
constexpr int BAR_PAIR = FT_NL + 32;     // FULL/EMPTY barriers: the 8 L warps + the one F warp that owns the buffer
enum { BAR_FULL0 = 1, BAR_EMPTY0 = 5, BAR_L = 9 };  // FULL0..3, EMPTY0..3

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// ---- Legendre helpers: CNT terms of one zonal wavenumber, polynomials Pm[n] of the current latitude ------------
template <int CNT>
__device__ __forceinline__ void leg_syn(const double (&x)[CNT], const double *__restrict__ Pm, double &e, double &o) {
    e = 0.0, o = 0.0;  // legendre.f90:148-161: even = sum over n = 1,3,.. ; odd = sum over n = 2,4,.. (1-based)
#pragma unroll
    for (int n = 0; n + 1 < CNT; n += 2) {
        const double2 p = *reinterpret_cast<const double2 *>(Pm + n);
        e += x[n] * p.x;
        o += x[n + 1] * p.y;
    }
    if (CNT & 1) e += x[CNT - 1] * Pm[CNT - 1];
}
template <int CNT>
__device__ __forceinline__ void leg_ana(double (&acc)[CNT], const double *__restrict__ Pm, const double ev, const double od) {
#pragma unroll
    for (int n = 0; n + 1 < CNT; n += 2) {  // legendre.f90:206-218: dot_product accumulates j = 1..24 in order
        const double2 p = *reinterpret_cast<const double2 *>(Pm + n);
        acc[n] += p.x * ev;
        acc[n + 1] += p.y * od;
    }
    if (CNT & 1) acc[CNT - 1] += Pm[CNT - 1] * ev;
}

// cooperative cp.async of the Legendre slice of latitude pair j (31 x 32 doubles = 496 16-byte chunks) by the L warps
__device__ __forceinline__ void pslice_fetch(double *ps, const double *__restrict__ cpolj, int j, int lt) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int q = lt + r * FT_NL;  // chunk
        if (q < MX * NX / 2) {
            const int m = q >> 4, n2 = (q & 15) * 2;
            cp_async16(ps + 2 * q, cpolj + ((size_t)m * IY + j) * NX + n2);
        }
    }
    cp_async_commit();
}

// ===================================================================================== inverse: spec -> grid
// Pipeline: L warps write latitude pair g into Fs[g % KF]; F warp (g % KF) transforms it on its own: stage A (7 items)
// -> its private exchange buffer -> stage B (8 identical items = item 0 with shifted pointers) -> HBM.
struct LdFs {
    const double *p;
    __device__ __forceinline__ double operator()(int r) const { return p[r * TILE]; }
};
struct StGridH {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int i, double v) const { p[i * TILE] = v * sc; }
};
constexpr int S2G_KF = 4;
constexpr int S2G_THREADS = FT_NL + 32 * S2G_KF;
constexpr size_t S2G_SMEM = (size_t)(S2G_KF * (FS_ELEMS + ES_ELEMS) + 2 * PS_ELEMS) * sizeof(double);

template <int LW>
__device__ __forceinline__ void s2g_L_warp(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                           double *Fs, double *Ps, const int lane) {
    // zonal wavenumbers of this warp and their number of terms inside the nsh2 mask (legendre.f90:68-77): 32 - m
    constexpr int MA = LW, MB = 30 - LW, MC = 15 - LW, MD = 15 + LW;
    constexpr int CA = 32 - MA, CB = 32 - MB, CC = 32 - MC, CD = (LW == 0) ? 1 : 32 - MD;
    const int cc = lane >> 4, mem = lane & 15, lt = LW * 32 + lane;
    const double *cpolj = c.G->cpolj;
    double xa[CA], xb[CB], xc[CC], xd[CD];
    int g = 0;  // global latitude-pair counter (continuous across work items: every barrier arrival is matched)
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int h = w & 1, t = (w >> 1) % c.ntiles, f = (w >> 1) / c.ntiles;
        const double *X = refp(c, t, descs[f].src, h * 16 + mem);
        pslice_fetch(Ps + (g & 1) * PS_ELEMS, cpolj, 0, lt);
#define LOADX(arr, M, CNT)                                                                          \
    _Pragma("unroll") for (int n = 0; n < CNT; n++) arr[n] = __ldg(X + ((size_t)(2 * (M) + cc) + (size_t)M2 * n) * TILE);
        LOADX(xa, MA, CA)
        LOADX(xb, MB, CB)
        LOADX(xc, MC, CC)
        if (LW != 0) { LOADX(xd, MD, CD) }
#undef LOADX
        cp_async_wait<0>();
        bar_sync(BAR_L, FT_NL);
#pragma unroll 1
        for (int it = 0; it < IY; it++, g++) {
            const int b = g & 1, k = g % S2G_KF;
            if (it + 1 < IY) pslice_fetch(Ps + (b ^ 1) * PS_ELEMS, cpolj, it + 1, lt);
            if (g >= S2G_KF) bar_sync(BAR_EMPTY0 + k, BAR_PAIR);  // F warp k has consumed Fs[k]
            double *F = Fs + k * FS_ELEMS + mem;
            const double *P = Ps + b * PS_ELEMS;
            double e, o;
#define SYN(arr, M, CNT)                                                          \
    leg_syn<CNT>(arr, P + (M) * NX, e, o);                                        \
    F[(2 * (M) + cc) * TILE] = e + o;      /* hemisphere 0: row il+1-j (legendre.f90:163) */ \
    F[(2 * (M) + cc) * TILE + 16] = e - o; /* hemisphere 1: row j                */
            SYN(xa, MA, CA)
            SYN(xb, MB, CB)
            SYN(xc, MC, CC)
            if (LW != 0) { SYN(xd, MD, CD) }
#undef SYN
            bar_arrive(BAR_FULL0 + k, BAR_PAIR);
            cp_async_wait<0>();
            bar_sync(BAR_L, FT_NL);  // next slice visible; everybody done with this one
        }
    }
}

__device__ __forceinline__ void s2g_F_warp(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                           const double *Fs, double *Es, const int fw, const int lane) {
    const int hemi = lane >> 4, mem = lane & 15;
    const double *F = Fs + fw * FS_ELEMS + lane;
    double *s = Es + fw * ES_ELEMS + lane;
    const LdFs ld{F};
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int h = w & 1, t = (w >> 1) % c.ntiles, f = (w >> 1) / c.ntiles;
        const InvDesc d = descs[f];
        double *G = scp(c, t, d.dst, h * 16 + mem);
#pragma unroll 1
        for (int it = fw; it < IY; it += S2G_KF) {  // 24 % KF == 0: pair `it` of every item lives in buffer it % KF
            const int row = hemi ? it : IL - 1 - it;
            bar_sync(BAR_FULL0 + fw, BAR_PAIR);
            fftb_A0(ld, s), fftb_A1(ld, s), fftb_A2(ld, s), fftb_A3(ld, s), fftb_A4(ld, s), fftb_A5(ld, s), fftb_A6(ld, s);
            __syncwarp();
            bar_arrive(BAR_EMPTY0 + fw, BAR_PAIR);
            double *Grow = G + (size_t)row * IX * TILE;
            const double sc = d.kcos == 1 ? 1.0 : c_T.cosgr[row];
#pragma unroll 1
            for (int kk = 0; kk < 8; kk++)  // stage-B item kk == item 0 on inputs 12kk.., outputs kk + 8q
                fftb_B0(s + 12 * kk * TILE, StGridH{Grow + kk * TILE, sc});
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(S2G_THREADS, 1) k_spec2grid_fused(const Ctx c, const InvDesc *__restrict__ descs, int nwork) {
    extern __shared__ double smem[];
    double *Fs = smem, *Es = Fs + S2G_KF * FS_ELEMS, *Ps = Es + S2G_KF * ES_ELEMS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < FT_L_WARPS) {
        switch (warp) {
            case 0: s2g_L_warp<0>(c, descs, nwork, Fs, Ps, lane); break;
            case 1: s2g_L_warp<1>(c, descs, nwork, Fs, Ps, lane); break;
            case 2: s2g_L_warp<2>(c, descs, nwork, Fs, Ps, lane); break;
            case 3: s2g_L_warp<3>(c, descs, nwork, Fs, Ps, lane); break;
            case 4: s2g_L_warp<4>(c, descs, nwork, Fs, Ps, lane); break;
            case 5: s2g_L_warp<5>(c, descs, nwork, Fs, Ps, lane); break;
            case 6: s2g_L_warp<6>(c, descs, nwork, Fs, Ps, lane); break;
            default: s2g_L_warp<7>(c, descs, nwork, Fs, Ps, lane); break;
        }
    } else {
        s2g_F_warp(c, descs, nwork, Fs, Es, warp - FT_L_WARPS, lane);
    }
}

// ===================================================================================== forward: grid -> spec
// F warp k: cp.async of the grid rows of its NEXT latitude pair into its private staging buffer, stage A (8 identical
// items, products fused into the reads) -> private exchange buffer -> stage B (7 items) -> Fs[k]; the L warps then
// accumulate the Gaussian quadrature into their register-resident spectral coefficients.
template <int MODE> struct LdStage {
    const double *a, *b;  // staged rows of the current latitude pair (shared memory, already lane-offset)
    double k0, sc;
    __device__ __forceinline__ double operator()(int i) const {
        if (MODE == FM_PLAIN) return a[i * TILE];
        if (MODE == FM_COS) return a[i * TILE] * sc;
        if (MODE == FM_KE) {
            const double u = a[i * TILE], v = b[i * TILE];
            return 0.5 * (u * u + v * v);
        }
        if (MODE == FM_FLUXT) return (-a[i * TILE] * (b[i * TILE] - k0)) * sc;
        return (-a[i * TILE] * b[i * TILE]) * sc;
    }
};
struct StFs {
    double *p;
    double scale;
    __device__ __forceinline__ void operator()(int r, double v) const { p[r * TILE] = v * scale; }
};
template <int MODE> struct G2S {
    static constexpr int NOPS = (MODE == FM_KE || MODE == FM_FLUXT || MODE == FM_FLUX) ? 2 : 1;
    static constexpr int KF = (NOPS == 2) ? 2 : 3;  // F warps = depth of the Fourier ring (shared-memory budget)
    static constexpr int THREADS = FT_NL + 32 * KF;
    static constexpr size_t SMEM = (size_t)(KF * (FS_ELEMS + ES_ELEMS + NOPS * GS_ELEMS) + 2 * PS_ELEMS) * sizeof(double);
};

template <int MODE>
__device__ __forceinline__ void g2s_F_warp(const Ctx &c, const FwdDesc *__restrict__ descs, const int nwork, double *Fs,
                                           double *Es, double *Gs, const int fw, const int lane) {
    constexpr int NOPS = G2S<MODE>::NOPS, KF = G2S<MODE>::KF;
    double *F = Fs + fw * FS_ELEMS + lane;
    double *s = Es + fw * ES_ELEMS + lane;
    double *Gw = Gs + (size_t)fw * NOPS * GS_ELEMS;
    // cp.async of the grid rows of latitude pair `it` (both hemispheres, 16 members, all operands) into Gw
    auto prefetch = [&](const double *A, const double *B, int it) {
#pragma unroll
        for (int op = 0; op < NOPS; op++) {
            const double *src = op ? B : A;
            double *dst = Gw + (size_t)op * GS_ELEMS;
#pragma unroll 4
            for (int r = 0; r < 48; r++) {  // 96 points x 2 hemispheres x 8 chunks of 16 bytes = 1536 chunks
                const int q = lane + r * 32;
                const int seg = q & 7, hm = (q >> 3) & 1, i = q >> 4;
                const int row = hm ? it : IL - 1 - it;
                cp_async16(dst + (i * TILE + hm * 16 + seg * 2), src + ((size_t)row * IX + i) * TILE + seg * 2);
            }
        }
        cp_async_commit();
    };
    int w = blockIdx.x;
    if (w >= nwork) return;
    FwdDesc d = descs[(w >> 1) / c.ntiles];
    const double *A = refp(c, (w >> 1) % c.ntiles, d.a, (w & 1) * 16);
    const double *B = (NOPS == 2) ? refp(c, (w >> 1) % c.ntiles, d.b, (w & 1) * 16) : nullptr;
    prefetch(A, B, fw);
    int g = fw;  // global pair counter of this warp's pairs (for the EMPTY handshake)
    for (; w < nwork; w += gridDim.x) {
#pragma unroll 1
        for (int it = fw; it < IY; it += KF, g += KF) {
            cp_async_wait<0>();
            __syncwarp();
            LdStage<MODE> ld;
            ld.a = Gw + lane;
            ld.b = (NOPS == 2) ? Gw + GS_ELEMS + lane : nullptr;
            ld.k0 = d.k0;
            {
                const int row = (lane >> 4) ? it : IL - 1 - it;
                ld.sc = (d.kcos == 3) ? c_T.cosgr2[row] : c_T.cosgr[row];
            }
#pragma unroll 1
            for (int kk = 0; kk < 8; kk++) {  // stage-A item kk == item 0 on inputs kk + 8q, outputs 12kk..
                LdStage<MODE> l2 = ld;
                l2.a += kk * TILE;
                if (NOPS == 2) l2.b += kk * TILE;
                fftf_A0(l2, s + 12 * kk * TILE);
            }
            __syncwarp();
            // staging buffer is free: fetch this warp's next pair (possibly of the next work item)
            {
                int nit = it + KF, nw = w;
                if (nit >= IY) nit = fw, nw = w + gridDim.x;
                if (nw < nwork) {
                    if (nw != w) {
                        d = descs[(nw >> 1) / c.ntiles];
                        A = refp(c, (nw >> 1) % c.ntiles, d.a, (nw & 1) * 16);
                        B = (NOPS == 2) ? refp(c, (nw >> 1) % c.ntiles, d.b, (nw & 1) * 16) : nullptr;
                    }
                    prefetch(A, B, nit);
                }
            }
            if (g >= KF) bar_sync(BAR_EMPTY0 + fw, BAR_PAIR);  // L warps have consumed Fs[fw]
            const StFs st{F, c_T.fc[3]};
            fftf_B0(s, st), F[TILE] = 0.0;  // Im(m=0) := 0 (fourier.f90:117)
            fftf_B1(s, st), fftf_B2(s, st), fftf_B3(s, st), fftf_B4(s, st), fftf_B5(s, st), fftf_B6(s, st);
            bar_arrive(BAR_FULL0 + fw, BAR_PAIR);
            __syncwarp();
        }
    }
}

template <int MODE, int LW>
__device__ __forceinline__ void g2s_L_warp(const Ctx &c, const FwdDesc *__restrict__ descs, const FwdOut *__restrict__ outs,
                                           const int nwork, const double *Fs, double *Ps, const int lane) {
    constexpr int KF = G2S<MODE>::KF;
    constexpr int MA = LW, MB = 30 - LW, MC = 15 - LW, MD = 15 + LW;
    // terms kept by the direct transform: n = 1..trunc+1 under the nsh2 mask (legendre.f90:206-218): min(31, 32 - m)
    constexpr int CA = (MA == 0) ? 31 : 32 - MA, CB = 32 - MB, CC = 32 - MC, CD = (LW == 0) ? 1 : 32 - MD;
    const int cc = lane >> 4, mem = lane & 15, lt = LW * 32 + lane;
    const double *cpolj = c.G->cpolj;
    double aa[CA], ab[CB], ac[CC], ad[CD];
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int h = w & 1, t = (w >> 1) % c.ntiles, f = (w >> 1) / c.ntiles;
        pslice_fetch(Ps + (g & 1) * PS_ELEMS, cpolj, 0, lt);
#pragma unroll
        for (int n = 0; n < CA; n++) aa[n] = 0.0;
#pragma unroll
        for (int n = 0; n < CB; n++) ab[n] = 0.0;
#pragma unroll
        for (int n = 0; n < CC; n++) ac[n] = 0.0;
#pragma unroll
        for (int n = 0; n < CD; n++) ad[n] = 0.0;
        cp_async_wait<0>();
        bar_sync(BAR_L, FT_NL);
#pragma unroll 1
        for (int it = 0; it < IY; it++, g++) {
            const int b = g & 1, k = it % KF;
            if (it + 1 < IY) pslice_fetch(Ps + (b ^ 1) * PS_ELEMS, cpolj, it + 1, lt);
            bar_sync(BAR_FULL0 + k, BAR_PAIR);
            const double *F = Fs + k * FS_ELEMS + mem;
            const double *P = Ps + b * PS_ELEMS;
            const double wt = c_T.wt[it];
#define ANA(arr, M, CNT)                                                                        \
    {                                                                                           \
        const double fn = F[(2 * (M) + cc) * TILE], fs = F[(2 * (M) + cc) * TILE + 16];         \
        leg_ana<CNT>(arr, P + (M) * NX, (fn + fs) * wt, (fn - fs) * wt); /* legendre.f90:196-197 */ \
    }
            ANA(aa, MA, CA)
            ANA(ab, MB, CB)
            ANA(ac, MC, CC)
            if (LW != 0) ANA(ad, MD, CD)
#undef ANA
            bar_arrive(BAR_EMPTY0 + k, BAR_PAIR);
            cp_async_wait<0>();
            bar_sync(BAR_L, FT_NL);
        }
        // write the spectral field: every (m, n) of this warp's wavenumbers, zeros outside the mask
        double *X = refp(c, t, outs[descs[f].fidx].dst, h * 16 + mem);
#define STOREX(arr, M, CNT)                                                                        \
    _Pragma("unroll") for (int n = 0; n < NX; n++)                                                \
        X[((size_t)(2 * (M) + cc) + (size_t)M2 * n) * TILE] = (n < CNT) ? arr[n < CNT ? n : 0] : 0.0;
        STOREX(aa, MA, CA)
        STOREX(ab, MB, CB)
        STOREX(ac, MC, CC)
        if (LW != 0) { STOREX(ad, MD, CD) }
#undef STOREX
    }
}

template <int MODE>
__global__ void __launch_bounds__(G2S<MODE>::THREADS, 1) k_grid2spec_fused(const Ctx c, const FwdDesc *__restrict__ descs,
                                                                           const FwdOut *__restrict__ outs, int nwork) {
    constexpr int KF = G2S<MODE>::KF, NOPS = G2S<MODE>::NOPS;
    extern __shared__ double smem[];
    double *Fs = smem, *Es = Fs + KF * FS_ELEMS, *Gs = Es + KF * ES_ELEMS, *Ps = Gs + KF * NOPS * GS_ELEMS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < FT_L_WARPS) {
        switch (warp) {
            case 0: g2s_L_warp<MODE, 0>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 1: g2s_L_warp<MODE, 1>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 2: g2s_L_warp<MODE, 2>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 3: g2s_L_warp<MODE, 3>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 4: g2s_L_warp<MODE, 4>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 5: g2s_L_warp<MODE, 5>(c, descs, outs, nwork, Fs, Ps, lane); break;
            case 6: g2s_L_warp<MODE, 6>(c, descs, outs, nwork, Fs, Ps, lane); break;
            default: g2s_L_warp<MODE, 7>(c, descs, outs, nwork, Fs, Ps, lane); break;
        }
    } else {
        g2s_F_warp<MODE>(c, descs, nwork, Fs, Es, Gs, warp - FT_L_WARPS, lane);
    }
}

// ------------------------------------------------------------------------------------------------- launchers
static int fused_grid(int nwork) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return nwork < sms ? nwork : sms;  // persistent: one CTA per SM
}
void launch_spec2grid_fused(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static bool init = false;
    if (!init) {
        cudaFuncSetAttribute(k_spec2grid_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2G_SMEM);
        init = true;
    }
    const int nwork = nf * c.ntiles * 2;
    k_spec2grid_fused<<<fused_grid(nwork), S2G_THREADS, S2G_SMEM, s>>>(c, d, nwork);
}
template <int MODE> static void launch_g2s_mode(cudaStream_t s, const Ctx &c, const FwdDesc *d, const FwdOut *o, int nf) {
    static bool init = false;
    if (!init) {
        cudaFuncSetAttribute(k_grid2spec_fused<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2S<MODE>::SMEM);
        init = true;
    }
    const int nwork = nf * c.ntiles * 2;
    k_grid2spec_fused<MODE><<<fused_grid(nwork), G2S<MODE>::THREADS, G2S<MODE>::SMEM, s>>>(c, d, o, nwork);
}
void launch_grid2spec_fused(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, const FwdOut *o, int nf) {
    if (!nf) return;
    switch (mode) {
        case FM_PLAIN: launch_g2s_mode<FM_PLAIN>(s, c, d, o, nf); break;
        case FM_COS: launch_g2s_mode<FM_COS>(s, c, d, o, nf); break;
        case FM_KE: launch_g2s_mode<FM_KE>(s, c, d, o, nf); break;
        case FM_FLUXT: launch_g2s_mode<FM_FLUXT>(s, c, d, o, nf); break;
        default: launch_g2s_mode<FM_FLUX>(s, c, d, o, nf); break;
    }
}

}  // namespace spdy
