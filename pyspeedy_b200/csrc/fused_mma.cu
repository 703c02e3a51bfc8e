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
//     A fragments come pre-swizzled (sign and nsh2 mask applied, zero-padded to a multiple of 4 terms) from
//     GlobTables::pq_inv, one 256-byte row per k-slice and quad.  B fragments are the spectral coefficients of the
//     8 members: each L warp keeps those of its 8 wavenumbers in registers for the whole work item (72 doubles),
//     so the coefficients cross the L1 once per field instead of once per latitude quad.
//   F warps: one thread = one (latitude row, member) line, whole-line FFT in registers (fft96_reg_gen.cuh), rows
//     read from the slot, 96 grid values stored to HBM.  No exchange buffer, no CTA-wide barrier.  F warps 0,1
//     (hemisphere 0 / 1 rows) take the even slots, F warps 2,3 the odd ones.
//   L -> F hand-over per slot through named barriers FULL/EMPTY (bar.arrive / bar.sync, 128 L + 64 F threads).
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
    __threadfence_block();
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
template <int M> __device__ __forceinline__ void mq_load_a(MqA<M> &f, const double *__restrict__ Aq) {
#pragma unroll
    for (int s = 0; s < MQ_KS(M); s++) f.a[s] = __ldg(Aq + (size_t)(MQ_KOFF(M) + s) * 32);
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
    const double *pq = c.G->pq_inv + lane;
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
            const double *Aq = pq + (size_t)jq * MQ_KTOT * 32;
            MqA<LW> a0;
            MqA<30 - LW> a1;
            MqA<LW + 4> a2;
            MqA<26 - LW> a3;
            MqA<LW + 8> a4;
            MqA<22 - LW> a5;
            MqA<LW + 12> a6;
            MqA<M7> a7;
            mq_load_a(a0, Aq), mq_load_a(a1, Aq), mq_load_a(a2, Aq), mq_load_a(a3, Aq);
            mq_load_a(a4, Aq), mq_load_a(a5, Aq), mq_load_a(a6, Aq);
            if (LW != 3) mq_load_a(a7, Aq);
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
        cudaFuncSetAttribute(k_spec2grid_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MQ_SMEM);
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma<<<nwork < sms ? nwork : sms, 256, MQ_SMEM, s>>>(c, d, nwork);
}

}  // namespace spdy
