// speedy-b200: fused spectral -> grid transform with the Legendre contraction on the FP64 tensor cores.
//
// Reference semantics: legendre.f90:130-168 (inverse Legendre), fourier.f90:63-88 + fftpack.f90 (inverse FFT).
//
// Why this shape: with lane = member a Legendre polynomial is a warp-uniform operand and
// every way of delivering it (LDG/LDS broadcast) costs L1/LSU data-pipe cycles per FMA, which made the first fused
// kernels (fused.cu) pipe-bound.  mma.sync.m8n8k4.f64 takes the polynomials as a per-lane A fragment instead: one
// coalesced 256-byte load feeds 256 FMAs, at the full FP64 rate of the part (tools/microbench/dmma.cu).
//
// Work item = (field, tile, member group): members 8*grp .. 8*grp+7.  CTA = 4 "L" warps + 4 "F" warps, persistent.
//   GEMM per zonal wavenumber m and latitude quad jq (latitudes j = 4jq .. 4jq+3 of the half grid):
//       C[(jl,hemi)][member] = sum_n A[(jl,hemi)][n] * X[n][member],   A = sgn(hemi,n) * P(m,n,j)
//     hemi 0 is grid row il-1-j (even + odd parity sums), hemi 1 is row j (even - odd): folding the sign into A turns
//     the N/S symmetry into the M dimension, so one C tile holds 4 latitudes x 2 hemispheres = 8 rows and a
//     shared-memory slot is 8 latitude rows x 62 Fourier rows x 8 members = 31.5 KB (six slots in flight).
//     A fragments come pre-swizzled (nsh2 mask applied, zero-padded to a multiple of 4 terms) from
//     GlobTables::pq_inv, one 128-byte row per k-slice and quad; the hemisphere sign is a per-lane mask.  B fragments are the spectral coefficients of the
//     8 members: each L warp keeps those of its 8 wavenumbers in registers for the whole work item (72 doubles),
//     so the coefficients cross the L1 once per field instead of once per latitude quad.
//   F warps: one thread = one (latitude row, member) line, whole-line FFT in registers (fft96_reg_gen.cuh), rows
//     read from the slot, 96 grid values stored to HBM.  No exchange buffer, no CTA-wide barrier.  F warps 0,1
//     (hemisphere 0 / 1 rows) take the even slots, F warps 2,3 the odd ones.
//   L -> F hand-over per slot through named barriers FULL/EMPTY (bar.arrive / bar.sync, 128 L + 64 F threads).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace spdy {

__host__ __device__ constexpr int MQ_KS(int m) { return (32 - m + 3) / 4; }                       // k-slices (4 terms each) of wavenumber m
__host__ __device__ constexpr int MQ_KOFF(int m) { int o = 0; for (int i = 0; i < m; i++) o += MQ_KS(i); return o; }
constexpr int MQ_KTOT = MQ_KOFF(MX);                                          // 143
static_assert(MQ_KTOT == PQ_KTOT, "GlobTables::pq_inv layout");
constexpr int MQ_NM = 8;                      // members per work item
constexpr int MQ_RS = M2 * MQ_NM + 8;         // slot row stride (doubles): 62 x 64 B + 64 B -> rows 2i, 2i+1 in
                                              // different bank halves: C stores and F row loads conflict-free
constexpr int MQ_SLOT = 8 * MQ_RS;            // doubles per slot
constexpr int MQ_NSLOT = 6;                   // even: slot parity = quad parity
constexpr size_t MQ_SMEM = (size_t)MQ_NSLOT * MQ_SLOT * sizeof(double);       // 193,536 bytes
constexpr int MQ_BARN = 128 + 64;             // 4 L warps + the 2 F warps of the slot's parity
enum { MQ_FULL0 = 1, MQ_EMPTY0 = 1 + MQ_NSLOT };  // named barriers 1..6 and 7..12

__device__ __forceinline__ void mq_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(MQ_BARN) : "memory"); }
__device__ __forceinline__ void mq_bar_arrive(int id) {
#ifdef MQ_FENCE
    __threadfence_block();
#endif
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(MQ_BARN) : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// B fragments (spectral coefficients) of wavenumber M: term n = 4s + lane%4, member lane/4; re and im parts.
template <int M> struct MqB {
    static constexpr int KS = MQ_KS(M);
    double br[KS], bi[KS];
};
template <int M> __device__ __forceinline__ void mq_load_b(MqB<M> &f, const double *__restrict__ Xl, const int kk) {
    constexpr int KS = MQ_KS(M), NMAX = 31 - M;
#pragma unroll
    for (int s = 0; s < KS; s++) {
        const double *x = Xl + (size_t)((2 * M) + M2 * 4 * s) * TILE;
        f.br[s] = __ldg(x), f.bi[s] = __ldg(x + TILE);
        if (4 * s + 3 > NMAX) {  // zero-padded terms: the table holds 0, keep the product finite whatever memory holds
            const bool in = (4 * s + kk <= NMAX);
            f.br[s] = in ? f.br[s] : 0.0, f.bi[s] = in ? f.bi[s] : 0.0;
        }
    }
}
// A fragments of wavenumber M for one latitude quad
template <int M> struct MqA {
    static constexpr int KS = MQ_KS(M);
    double a[KS];
};
// The hemisphere-1 half of a fragment is the hemisphere-0 half with the sign of the odd terms flipped, and the
// parity of n = 4s + lane%4 is a lane constant: the table holds 16 values per k-slice (18 KB per quad, L1-resident
// next to 190 KB of shared memory) and the lane applies its sign mask.
template <int M> __device__ __forceinline__ void mq_load_a(MqA<M> &f, const double *__restrict__ Aq, const long long sgn) {
#pragma unroll
    for (int s = 0; s < MQ_KS(M); s++)
        f.a[s] = __longlong_as_double(__double_as_longlong(__ldg(Aq + (size_t)(MQ_KOFF(M) + s) * 16)) ^ sgn);
}
//   Sl : slot + lane offset (row L/4, members 2*(L%4), 2*(L%4)+1)
template <int M> __device__ __forceinline__ void mq_mma_store(const MqA<M> &fa, const MqB<M> &fb, double *__restrict__ Sl) {
    double r0 = 0.0, r1 = 0.0, i0 = 0.0, i1 = 0.0;
#pragma unroll
    for (int s = 0; s < MQ_KS(M); s++) {
        dmma884(r0, r1, fa.a[s], fb.br[s]);
        dmma884(i0, i1, fa.a[s], fb.bi[s]);
    }
    *reinterpret_cast<double2 *>(Sl + (2 * M) * MQ_NM) = make_double2(r0, r1);
    *reinterpret_cast<double2 *>(Sl + (2 * M + 1) * MQ_NM) = make_double2(i0, i1);
}

// two wavenumbers at once: four independent DMMA chains cover the tensor-pipe latency
template <int MA, int MB>
__device__ __forceinline__ void mq_mma_store2(const MqA<MA> &fa, const MqB<MA> &fb, const MqA<MB> &ga, const MqB<MB> &gb,
                                              double *__restrict__ Sl) {
    constexpr int KA = MQ_KS(MA), KB = MQ_KS(MB), KMAX = KA > KB ? KA : KB;
    double r0 = 0.0, r1 = 0.0, i0 = 0.0, i1 = 0.0, p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        if (s < KA) dmma884(r0, r1, fa.a[s], fb.br[s]);
        if (s < KB) dmma884(p0, p1, ga.a[s], gb.br[s]);
        if (s < KA) dmma884(i0, i1, fa.a[s], fb.bi[s]);
        if (s < KB) dmma884(q0, q1, ga.a[s], gb.bi[s]);
    }
    *reinterpret_cast<double2 *>(Sl + (2 * MA) * MQ_NM) = make_double2(r0, r1);
    *reinterpret_cast<double2 *>(Sl + (2 * MA + 1) * MQ_NM) = make_double2(i0, i1);
    *reinterpret_cast<double2 *>(Sl + (2 * MB) * MQ_NM) = make_double2(p0, p1);
    *reinterpret_cast<double2 *>(Sl + (2 * MB + 1) * MQ_NM) = make_double2(q0, q1);
}

// L warp LW owns the units u = LW, LW+4, LW+8, LW+12 (wavenumber pairs (u, 30-u); u = 15 is m = 15 alone): 35-36
// k-slices per quad for every warp.  The A fragments of wavenumber i+1 are requested before the DMMAs of wavenumber
// i are issued (explicit software pipeline: with one L warp per scheduler nothing else hides the L2 latency), and
// the first request of a quad is issued before the slot's EMPTY barrier.
template <int LW>
__device__ __forceinline__ void s2g_mma_L(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork, double *slots,
                                          const int lane) {
    constexpr int M7 = (LW != 3) ? 18 - LW : 15;  // warp 3 has only seven wavenumbers (slot M7 unused there)
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_inv + (lane & 15);
    const long long sgn = (lane >= 16 && (lane & 1)) ? (long long)0x8000000000000000ull : 0ll;
    int g = 0;  // quad counter, continuous across work items
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const double *Xl = refp(c, t, descs[f].src, 0) + (size_t)(M2 * kk) * TILE + MQ_NM * grp + col;
        MqB<LW> b0;
        MqB<30 - LW> b1;
        MqB<LW + 4> b2;
        MqB<26 - LW> b3;
        MqB<LW + 8> b4;
        MqB<22 - LW> b5;
        MqB<LW + 12> b6;
        MqB<M7> b7;
        mq_load_b(b0, Xl, kk), mq_load_b(b1, Xl, kk), mq_load_b(b2, Xl, kk), mq_load_b(b3, Xl, kk);
        mq_load_b(b4, Xl, kk), mq_load_b(b5, Xl, kk), mq_load_b(b6, Xl, kk);
        if (LW != 3) mq_load_b(b7, Xl, kk);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % MQ_NSLOT;
            double *Sl = slots + sl * MQ_SLOT + col * MQ_RS + 2 * kk;
            const double *Aq = pq + (size_t)jq * MQ_KTOT * 16;
            MqA<LW> a0;
            MqA<30 - LW> a1;
            MqA<LW + 4> a2;
            MqA<26 - LW> a3;
            MqA<LW + 8> a4;
            MqA<22 - LW> a5;
            MqA<LW + 12> a6;
            MqA<M7> a7;
            mq_load_a(a0, Aq, sgn), mq_load_a(a1, Aq, sgn), mq_load_a(a2, Aq, sgn), mq_load_a(a3, Aq, sgn);
            mq_load_a(a4, Aq, sgn), mq_load_a(a5, Aq, sgn), mq_load_a(a6, Aq, sgn);
            if (LW != 3) mq_load_a(a7, Aq, sgn);
            if (g >= MQ_NSLOT) mq_bar_sync(MQ_EMPTY0 + sl);
            mq_mma_store2(a0, b0, a1, b1, Sl);
            mq_mma_store2(a2, b2, a3, b3, Sl);
            mq_mma_store2(a4, b4, a5, b5, Sl);
            if (LW != 3) mq_mma_store2(a6, b6, a7, b7, Sl);
            else mq_mma_store(a6, b6, Sl);
            mq_bar_arrive(MQ_FULL0 + sl);
        }
    }
}

struct LdSlot {
    const double *p;
    __device__ __forceinline__ double operator()(int r) const { return p[r * MQ_NM]; }
};
struct StGridH2 {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int i, double v) const { p[i * TILE] = v * sc; }
};

// F warp fw: hemisphere fw & 1 (slot rows 4*hemi .. 4*hemi+3), slots of parity fw >> 1; lane = (row in 0..3, member)
__device__ __forceinline__ void s2g_mma_F(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                          const double *slots, const int fw, const int lane) {
    const int hemi = fw & 1, par = fw >> 1, jl = lane >> 3, mem = lane & 7, row = 4 * hemi + jl;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const InvDesc d = descs[f];
        double *G = scp(c, t, d.dst, MQ_NM * grp + mem);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            if ((g & 1) != par) continue;
            const int sl = g % MQ_NSLOT;
            const int j = 4 * jq + jl, lat = hemi ? j : IL - 1 - j;  // legendre.f90:163-167
            const LdSlot ld{slots + sl * MQ_SLOT + row * MQ_RS + mem};
            const StGridH2 st{G + (size_t)lat * IX * TILE, d.kcos == 1 ? 1.0 : c_T.cosgr[lat]};
            double x[IX];
            mq_bar_sync(MQ_FULL0 + sl);
            rfftb_A0(ld, x), rfftb_A1(ld, x), rfftb_A2(ld, x), rfftb_A3(ld, x), rfftb_A4(ld, x), rfftb_A5(ld, x),
                rfftb_A6(ld, x);
            mq_bar_arrive(MQ_EMPTY0 + sl);  // all 62 rows of this thread's line are in registers
            rfftb_B0(x, st), rfftb_B1(x, st), rfftb_B2(x, st), rfftb_B3(x, st), rfftb_B4(x, st), rfftb_B5(x, st),
                rfftb_B6(x, st), rfftb_B7(x, st);
        }
    }
}

__global__ void __launch_bounds__(256, 1) k_spec2grid_mma(const Ctx c, const InvDesc *__restrict__ descs, int nwork) {
    extern __shared__ __align__(16) double mq_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    switch (warp) {
        case 0: s2g_mma_L<0>(c, descs, nwork, mq_sm, lane); break;
        case 1: s2g_mma_L<1>(c, descs, nwork, mq_sm, lane); break;
        case 2: s2g_mma_L<2>(c, descs, nwork, mq_sm, lane); break;
        case 3: s2g_mma_L<3>(c, descs, nwork, mq_sm, lane); break;
        default: s2g_mma_F(c, descs, nwork, mq_sm, warp - 4, lane); break;
    }
}

void launch_spec2grid_mma(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_spec2grid_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MQ_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_spec2grid_mma needs %zu bytes of shared memory per CTA (sm_100a)\n", MQ_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma<<<nwork < sms ? nwork : sms, 256, MQ_SMEM, s>>>(c, d, nwork);
}


// =================================================================================== forward: grid -> spectral
// Reference semantics: fourier.f90:90-123 (+ fftpack.f90) and legendre.f90:170-221.
//
// The mirror image of the kernel above.  F warps: one thread = one (latitude row, member) line; the grid-point
// products of tendencies.f90:238-268 are applied while loading (LdGrid<MODE>, transforms.cu), whole-line forward
// FFT in registers, the 62 Fourier rows go to the slot.  L warps: the Gaussian quadrature as DMMA,
//     X[n][member] += A[n][(jl,hemi)] * F[(jl,hemi)][member],  A = sgn(hemi,n) * wt(j) * P(m,n,j)
// with K = the 8 latitude rows of a slot (two k-slices: hemisphere 0 rows, hemisphere 1 rows), M = 8 spectral rows n,
// accumulators C[n-tile][8 members] kept in registers over the six latitude quads of a work item (20 tiles x {re,im}
// per L warp = 160 registers) and stored once.  A fragments pre-swizzled on the host (GlobTables::pq_dir).
__host__ __device__ constexpr int MD_NT(int m) { return ((30 < 31 - m ? 30 : 31 - m) + 1 + 7) / 8; }  // 8-row n-tiles
__host__ __device__ constexpr int MD_TOFF(int m) { int o = 0; for (int i = 0; i < m; i++) o += MD_NT(i); return o; }
constexpr int MD_TTOT = MD_TOFF(MX);  // 79
static_assert(MD_TTOT == PD_TTOT, "GlobTables::pq_dir layout");
constexpr int MD_RS = M2 * MQ_NM + 4;         // slot row stride: 62 x 64 B + 32 B -> the four rows of a B fragment
                                              // fall into different bank quarters (conflict-free LDS.64)
constexpr int MD_SLOT = 8 * MD_RS;
constexpr int MD_NSLOT = 6;
constexpr size_t MD_SMEM = (size_t)MD_NSLOT * MD_SLOT * sizeof(double);  // 192,000 bytes
enum { MD_FULL0 = 1, MD_EMPTY0 = 1 + MD_NSLOT };

template <int M> struct MdC {  // accumulators of wavenumber M: [n-tile][c0, c1], real and imaginary part
    static constexpr int NT = MD_NT(M);
    double cr[NT][2], ci[NT][2];
};
template <int M> __device__ __forceinline__ void md_zero(MdC<M> &c) {
#pragma unroll
    for (int i = 0; i < MD_NT(M); i++) c.cr[i][0] = c.cr[i][1] = c.ci[i][0] = c.ci[i][1] = 0.0;
}
//   Aq : fragment table of this quad + lane ; Bl : slot + lane offset (row L%4, member L/4)
template <int M>
__device__ __forceinline__ void md_mma(MdC<M> &c, const double *__restrict__ Aq, const double *__restrict__ Bl) {
    constexpr int NT = MD_NT(M);
    double a0[NT], a1[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) {
        a0[i] = __ldg(Aq + (size_t)((MD_TOFF(M) + i) * 2) * 32);
        a1[i] = __ldg(Aq + (size_t)((MD_TOFF(M) + i) * 2 + 1) * 32);
    }
    const double br0 = Bl[(2 * M) * MQ_NM], br1 = Bl[4 * MD_RS + (2 * M) * MQ_NM];
    const double bi0 = Bl[(2 * M + 1) * MQ_NM], bi1 = Bl[4 * MD_RS + (2 * M + 1) * MQ_NM];
#pragma unroll
    for (int i = 0; i < NT; i++) {
        dmma884(c.cr[i][0], c.cr[i][1], a0[i], br0);
        dmma884(c.ci[i][0], c.ci[i][1], a0[i], bi0);
    }
#pragma unroll
    for (int i = 0; i < NT; i++) {
        dmma884(c.cr[i][0], c.cr[i][1], a1[i], br1);
        dmma884(c.ci[i][0], c.ci[i][1], a1[i], bi1);
    }
}
//   Xl : spectral output + lane offset (row L/4 of an n-tile, members 2*(L%4), +1); rows up to n = 31 are written
template <int M> __device__ __forceinline__ void md_store(const MdC<M> &c, double *__restrict__ Xl) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double2 *pr = reinterpret_cast<double2 *>(Xl + (size_t)((2 * M) + M2 * 8 * i) * TILE);
        double2 *pi = reinterpret_cast<double2 *>(Xl + (size_t)((2 * M + 1) + M2 * 8 * i) * TILE);
        if (i < MD_NT(M)) {
            *pr = make_double2(c.cr[i < MD_NT(M) ? i : 0][0], c.cr[i < MD_NT(M) ? i : 0][1]);
            *pi = make_double2(c.ci[i < MD_NT(M) ? i : 0][0], c.ci[i < MD_NT(M) ? i : 0][1]);
        } else {  // legendre.f90:206-218: rows outside the nsh2 mask are zero
            *pr = make_double2(0.0, 0.0), *pi = make_double2(0.0, 0.0);
        }
    }
}

template <int LW>
__device__ __forceinline__ void g2s_mma_L(const Ctx &c, const FwdDesc *__restrict__ descs, const FwdOut *__restrict__ outs,
                                          const int nwork, const double *slots, const int lane) {
    constexpr int M7 = (LW != 3) ? 18 - LW : 15;
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_dir + lane;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        MdC<LW> c0;
        MdC<30 - LW> c1;
        MdC<LW + 4> c2;
        MdC<26 - LW> c3;
        MdC<LW + 8> c4;
        MdC<22 - LW> c5;
        MdC<LW + 12> c6;
        MdC<M7> c7;
        md_zero(c0), md_zero(c1), md_zero(c2), md_zero(c3), md_zero(c4), md_zero(c5), md_zero(c6), md_zero(c7);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % MD_NSLOT;
            const double *Bl = slots + sl * MD_SLOT + kk * MD_RS + col;
            const double *Aq = pq + (size_t)jq * (MD_TTOT * 2 * 32);
            asm volatile("bar.sync %0, %1;" ::"r"(MD_FULL0 + sl), "n"(MQ_BARN) : "memory");
            md_mma(c0, Aq, Bl), md_mma(c1, Aq, Bl), md_mma(c2, Aq, Bl), md_mma(c3, Aq, Bl);
            md_mma(c4, Aq, Bl), md_mma(c5, Aq, Bl), md_mma(c6, Aq, Bl);
            if (LW != 3) md_mma(c7, Aq, Bl);
            asm volatile("bar.arrive %0, %1;" ::"r"(MD_EMPTY0 + sl), "n"(MQ_BARN) : "memory");
        }
        double *Xl = refp(c, t, outs[descs[f].fidx].dst, 0) + (size_t)(M2 * col) * TILE + MQ_NM * grp + 2 * kk;
        md_store(c0, Xl), md_store(c1, Xl), md_store(c2, Xl), md_store(c3, Xl);
        md_store(c4, Xl), md_store(c5, Xl), md_store(c6, Xl);
        if (LW != 3) md_store(c7, Xl);
    }
}

struct StSlot {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int r, double v) const { p[r * MQ_NM] = v * sc; }
};

template <int MODE>
__device__ __forceinline__ void g2s_mma_F(const Ctx &c, const FwdDesc *__restrict__ descs, const int nwork, double *slots,
                                          const int fw, const int lane) {
    const int hemi = fw & 1, par = fw >> 1, jl = lane >> 3, mem = lane & 7, row = 4 * hemi + jl;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const FwdDesc d = descs[f];
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            if ((g & 1) != par) continue;
            const int sl = g % MD_NSLOT;
            const int j = 4 * jq + jl, lat = hemi ? j : IL - 1 - j;  // legendre.f90:196-197: fn = row il+1-j, fs = row j
            const LdGrid<MODE> ld = make_ld<MODE>(c, t, d, lat, MQ_NM * grp + mem);
            {   // this warp's next task (two quads further on, possibly in its next work item): pull its grid rows
                // into L2 now -- the whole-line FFT leaves no registers to prefetch into
                int nq = jq + 2, nw = w;
                if (nq >= IY / 4) nq -= IY / 4, nw += gridDim.x;
                if (nw < nwork) {
                    const int ngrp = nw & 3, nt = (nw >> 2) % c.ntiles, nf = (nw >> 2) / c.ntiles;
                    const int nj = 4 * nq + jl, nlat = hemi ? nj : IL - 1 - nj;
                    const LdGrid<MODE> ln = make_ld<MODE>(c, nt, nw == w ? d : descs[nf], nlat, MQ_NM * ngrp + mem);
#pragma unroll 8
                    for (int i = 0; i < IX; i++) {
                        prefetch_l2(ln.a + i * TILE);
                        if (MODE == FM_KE || MODE == FM_FLUXT || MODE == FM_FLUX) prefetch_l2(ln.b + i * TILE);
                    }
                }
            }
            double x[IX];
            rfftf_A0(ld, x), rfftf_A1(ld, x), rfftf_A2(ld, x), rfftf_A3(ld, x), rfftf_A4(ld, x), rfftf_A5(ld, x),
                rfftf_A6(ld, x), rfftf_A7(ld, x);
            if (g >= MD_NSLOT) asm volatile("bar.sync %0, %1;" ::"r"(MD_EMPTY0 + sl), "n"(MQ_BARN) : "memory");
            double *S = slots + sl * MD_SLOT + row * MD_RS + mem;
            const StSlot st{S, c_T.fc[3]};
            rfftf_B0(x, st), rfftf_B1(x, st), rfftf_B2(x, st), rfftf_B3(x, st), rfftf_B4(x, st), rfftf_B5(x, st),
                rfftf_B6(x, st);
            S[MQ_NM] = 0.0;  // fourier.f90:117: Im of m = 0
            __threadfence_block();
            asm volatile("bar.arrive %0, %1;" ::"r"(MD_FULL0 + sl), "n"(MQ_BARN) : "memory");
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_grid2spec_mma(const Ctx c, const FwdDesc *__restrict__ descs,
                                                          const FwdOut *__restrict__ outs, int nwork) {
    extern __shared__ __align__(16) double md_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    switch (warp) {
        case 0: g2s_mma_L<0>(c, descs, outs, nwork, md_sm, lane); break;
        case 1: g2s_mma_L<1>(c, descs, outs, nwork, md_sm, lane); break;
        case 2: g2s_mma_L<2>(c, descs, outs, nwork, md_sm, lane); break;
        case 3: g2s_mma_L<3>(c, descs, outs, nwork, md_sm, lane); break;
        default: g2s_mma_F<MODE>(c, descs, nwork, md_sm, warp - 4, lane); break;
    }
}

template <int MODE> static void launch_g2s_mma_mode(cudaStream_t s, const Ctx &c, const FwdDesc *d, const FwdOut *o, int nf) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_grid2spec_mma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_grid2spec_mma needs %zu bytes of shared memory per CTA (sm_100a)\n", MD_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_grid2spec_mma<MODE><<<nwork < sms ? nwork : sms, 256, MD_SMEM, s>>>(c, d, o, nwork);
}
void launch_grid2spec_mma(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, const FwdOut *o, int nf) {
    if (!nf) return;
    switch (mode) {
        case FM_PLAIN: launch_g2s_mma_mode<FM_PLAIN>(s, c, d, o, nf); break;
        case FM_COS: launch_g2s_mma_mode<FM_COS>(s, c, d, o, nf); break;
        case FM_KE: launch_g2s_mma_mode<FM_KE>(s, c, d, o, nf); break;
        case FM_FLUXT: launch_g2s_mma_mode<FM_FLUXT>(s, c, d, o, nf); break;
        default: launch_g2s_mma_mode<FM_FLUX>(s, c, d, o, nf); break;
    }
}

}  // namespace spdy
