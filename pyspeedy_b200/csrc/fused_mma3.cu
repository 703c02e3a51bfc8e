// speedy-b200: third-generation fused spectral -> grid transform (parity-pure DMMA over latitude octets).
//
// Reference semantics: legendre.f90:130-168 (inverse Legendre), fourier.f90:63-88 + fftpack.f90:69-134 (inverse FFT).
//
// k_spec2grid_mma2 (fused_mma2.cu) folds the N/S symmetry into the M dimension of the DMMA (4 latitudes x 2
// hemispheres), which computes every product twice.  Measured on the forward kernel: every DMMA removed is 16 cycles
// less on the critical path, because DMMA and the latency-bound FFT butterflies share the FP64 pipe.  Here the M
// dimension is EIGHT latitudes and a k-slice holds four n-terms of ONE parity:
//     E[j][member] = sum over even n of P(m,n,j) X[n],   O[j][member] = sum over odd n,
//     row il-1-j = E + O,  row j = E - O          (legendre.f90:163-167)
// 155 k-slices per octet instead of 2 x 143 per two quads: 310 DMMAs per 8 latitudes instead of 572.
// A slot is an octet (16 latitude rows x 62 x 8 members = 63 KB), two slots; the FFT side is that of k_spec2grid_mma2
// (four warps per hemisphere, two-stage FFT, in-place stage B, one TMA tensor store per pass through the same 5-D
// tensor map) taking a slot in two passes per hemisphere (4 latitude rows x 8 members each).  224 KB of shared memory.
// Measured (512 members, 77 fields): 0.469 ms against 0.496 for k_spec2grid_mma2.  A variant with all eight FFT warps
// on one pass (one item per warp and stage, one exchange double buffer, 175 KB) was slower (0.621 ms: twice the
// barriers per latitude, and its eight-way item switch cost 1.8 KB of spills in the Legendre warps).
#include "fused_common.cuh"

namespace spdy {

__host__ __device__ constexpr int P3_KSE(int m) { return ((32 - m + 1) / 2 + 3) / 4; }  // k-slices of even n (n = 0, 2, ..)
__host__ __device__ constexpr int P3_KSO(int m) { return ((32 - m) / 2 + 3) / 4; }      // k-slices of odd n
__host__ __device__ constexpr int P3_KOFF(int m) { int o = 0; for (int i = 0; i < m; i++) o += P3_KSE(i) + P3_KSO(i); return o; }
static_assert(P3_KOFF(MX) == PQ2_KTOT, "GlobTables::pq_inv2 layout");
constexpr int P3_SLOT = 16 * MQ_RS;                                   // doubles per slot (octet)
constexpr size_t P3_SMEM = ((size_t)2 * P3_SLOT + 4 * M2_XH) * sizeof(double);
enum { P3_FULL0 = 1, P3_EMPTY0 = 3, P3_GRP0 = 5 };  // + 4 group barriers
static_assert(P3_SMEM <= 232448, "shared memory per CTA on sm_100a");

// B fragments (spectral coefficients) of wavenumber M, parity-pure: term n = par + 2 * (4s + lane%4), member lane/4
template <int M> struct P3B {
    double er[P3_KSE(M)], ei[P3_KSE(M)], orr[P3_KSO(M)], oi[P3_KSO(M)];
};
template <int M> __device__ __forceinline__ void p3_load_b(P3B<M> &f, const double *__restrict__ Xl, const int kk) {
    constexpr int NMAX = 31 - M;
#pragma unroll
    for (int s = 0; s < P3_KSE(M); s++) {
        const int n = 2 * (4 * s + kk);
        const bool in = (n <= NMAX);  // zero-padded terms: the table holds 0, keep the product finite whatever memory holds
        const double *x = Xl + (size_t)((2 * M) + M2 * (in ? n : 0)) * TILE;
        const double r = __ldg(x), i = __ldg(x + TILE);
        f.er[s] = in ? r : 0.0, f.ei[s] = in ? i : 0.0;
    }
#pragma unroll
    for (int s = 0; s < P3_KSO(M); s++) {
        const int n = 1 + 2 * (4 * s + kk);
        const bool in = (n <= NMAX);
        const double *x = Xl + (size_t)((2 * M) + M2 * (in ? n : 0)) * TILE;
        const double r = __ldg(x), i = __ldg(x + TILE);
        f.orr[s] = in ? r : 0.0, f.oi[s] = in ? i : 0.0;
    }
}
template <int M> __device__ __forceinline__ void p3_prefetch_b(const double *__restrict__ Xl, const int kk) {
    constexpr int NMAX = 31 - M;
#pragma unroll
    for (int s = 0; s < P3_KSE(M) + P3_KSO(M); s++) {
        const int n = (s < P3_KSE(M)) ? 2 * (4 * s + kk) : 1 + 2 * (4 * (s - P3_KSE(M)) + kk);
        if (n <= NMAX) {
            const double *x = Xl + (size_t)((2 * M) + M2 * n) * TILE;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(x));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(x + TILE));
        }
    }
}
// A fragments of wavenumber M for one octet, requested one wavenumber ahead of their use
template <int M> struct P3A {
    double ae[P3_KSE(M)], ao[P3_KSO(M)];
};
//   Aq : fragment table of this octet + lane
template <int M> __device__ __forceinline__ void p3_load_a(P3A<M> &f, const double *__restrict__ Aq) {
#pragma unroll
    for (int s = 0; s < P3_KSE(M); s++) f.ae[s] = __ldg(Aq + (size_t)(P3_KOFF(M) + s) * 32);
#pragma unroll
    for (int s = 0; s < P3_KSO(M); s++) f.ao[s] = __ldg(Aq + (size_t)(P3_KOFF(M) + P3_KSE(M) + s) * 32);
}
//   Sl : slot + lane offset (row L/4, members 2*(L%4), 2*(L%4)+1)
template <int M> __device__ __forceinline__ void p3_mma_store(const P3B<M> &fb, const P3A<M> &fa, double *__restrict__ Sl) {
    constexpr int KE = P3_KSE(M), KO = P3_KSO(M);
    double er0 = 0.0, er1 = 0.0, ei0 = 0.0, ei1 = 0.0, or0 = 0.0, or1 = 0.0, oi0 = 0.0, oi1 = 0.0;
#pragma unroll
    for (int s = 0; s < KE; s++) {
        dmma884(er0, er1, fa.ae[s], fb.er[s]);
        dmma884(ei0, ei1, fa.ae[s], fb.ei[s]);
        if (s < KO) {
            dmma884(or0, or1, fa.ao[s], fb.orr[s]);
            dmma884(oi0, oi1, fa.ao[s], fb.oi[s]);
        }
    }
    // rows 0..7 of the slot: latitude il-1-j (even + odd), rows 8..15: latitude j (even - odd)
    *reinterpret_cast<double2 *>(Sl + (2 * M) * MQ_NM) = make_double2(er0 + or0, er1 + or1);
    *reinterpret_cast<double2 *>(Sl + (2 * M + 1) * MQ_NM) = make_double2(ei0 + oi0, ei1 + oi1);
    *reinterpret_cast<double2 *>(Sl + 8 * MQ_RS + (2 * M) * MQ_NM) = make_double2(er0 - or0, er1 - or1);
    *reinterpret_cast<double2 *>(Sl + 8 * MQ_RS + (2 * M + 1) * MQ_NM) = make_double2(ei0 - oi0, ei1 - oi1);
}

// L warp LW of 8: wavenumbers LW, 30-LW, LW+8, 22-LW (LW = 7: 7, 23, 15): 20 k-slices per warp (15 for LW = 7)
template <int LW>
__device__ __forceinline__ void s2g3_L(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork, double *slots,
                                       const int lane) {
    constexpr int M3 = (LW != 7) ? 22 - LW : 15;
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_inv2 + lane;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const double *Xl = refp(c, t, descs[f].src, 0) + MQ_NM * grp + col;
        P3B<LW> b0;
        P3B<30 - LW> b1;
        P3B<LW + 8> b2;
        P3B<M3> b3;
        p3_load_b(b0, Xl, kk), p3_load_b(b1, Xl, kk), p3_load_b(b2, Xl, kk);
        if (LW != 7) p3_load_b(b3, Xl, kk);
        if (w + (int)gridDim.x < nwork) {  // the coefficients of this warp's next work item: pull them into L2 now
            const int wn = w + gridDim.x, tn = (wn >> 2) % c.ntiles, fn = (wn >> 2) / c.ntiles;
            const double *Xn = refp(c, tn, descs[fn].src, 0) + MQ_NM * (wn & 3) + col;
            p3_prefetch_b<LW>(Xn, kk), p3_prefetch_b<30 - LW>(Xn, kk), p3_prefetch_b<LW + 8>(Xn, kk);
            if (LW != 7) p3_prefetch_b<M3>(Xn, kk);
        }
#pragma unroll 1
        for (int jo = 0; jo < IY / 8; jo++, g++) {
            const int sl = g & 1;
            double *Sl = slots + sl * P3_SLOT + col * MQ_RS + 2 * kk;
            const double *Aq = pq + (size_t)jo * PQ2_KTOT * 32;
            P3A<LW> a0;
            P3A<30 - LW> a1;
            P3A<LW + 8> a2;
            P3A<M3> a3;
            p3_load_a(a0, Aq);
            if (g >= 2) m2_sync(P3_EMPTY0 + sl, 512);
            p3_load_a(a1, Aq);
            p3_mma_store(b0, a0, Sl);
            p3_load_a(a2, Aq);
            p3_mma_store(b1, a1, Sl);
            if (LW != 7) p3_load_a(a3, Aq);
            p3_mma_store(b2, a2, Sl);
            if (LW != 7) p3_mma_store(b3, a3, Sl);
            m2_arrive(P3_FULL0 + sl, 512);
        }
    }
}

struct StExch1 {  // StExchK (fused_common.cuh) without the scaling
    double *p;
    __device__ __forceinline__ void operator()(int i, double v) const { p[(i >> 3) * 32] = v; }
};

// F warp fw of 8: hemisphere fw >> 2, item share fw & 3; lane = (jl, member); two passes (halves of the hemisphere's
// eight latitudes) per octet.  Hemisphere 1 rows 8 + 4*half + jl hold latitude 8jo + 4*half + jl; hemisphere 0 is read in
// reverse (slot row 4*half + 3 - jl = latitude il-4-8jo-4*half + jl) so that the four latitudes of a pass ascend with jl:
// one TMA box.  Same two-stage FFT, in-place stage B and tensor store as s2g2_F (fused_mma2.cu).
__device__ __forceinline__ void s2g3_F(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                       const double *slots, double *exch, const CUtensorMap *tmap, const int fw,
                                       const int lane) {
    const int hemi = fw >> 2, wq = fw & 3, jl = lane >> 3, mem = lane & 7;
    const bool issuer = (wq == 0 && lane == 0);
    int g = 0, p = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const InvDesc d = descs[f];
        const int ebase = (int)((long long)t * c.scr_elems + d.dst);
#pragma unroll 1
        for (int jo = 0; jo < IY / 8; jo++, g++) {
            const int sl = g & 1;
            m2_sync(P3_FULL0 + sl, 512);
#pragma unroll 1
            for (int half = 0; half < 2; half++, p++) {
                const int row = hemi ? 8 + 4 * half + jl : 4 * half + 3 - jl;
                const int lat0 = hemi ? 8 * jo + 4 * half : IL - 4 - 8 * jo - 4 * half, lat = lat0 + jl;
                const LdSlot ld{slots + sl * P3_SLOT + row * MQ_RS + mem};
                double *xbuf = exch + (size_t)(2 * hemi + (p & 1)) * M2_XH, *xb = xbuf + lane;
                if (wq == 0) {
                    fftb_A1(ld, xb), fftb_A0(ld, xb);
                } else if (wq == 1) {
                    fftb_A2(ld, xb), fftb_A6(ld, xb);
                } else if (wq == 2) {
                    fftb_A3(ld, xb), fftb_A4(ld, xb);
                } else {
                    fftb_A5(ld, xb);
                }
                if (half == 1) m2_arrive(P3_EMPTY0 + sl, 512);  // this warp has read its share of the slot completely
                m2_sync(P3_GRP0 + 2 * hemi, 128);
                if (d.kcos == 1) {  // 59 of the 77 fields: no 1/cos(lat) factor, no multiply per grid point
#pragma unroll 1
                    for (int k = 2 * wq; k < 2 * wq + 2; k++) fftb_B0(xb + 12 * k * 32, StExch1{xb + 12 * k * 32});
                } else {
                    const double sc = c_T.cosgr[lat];
#pragma unroll 1
                    for (int k = 2 * wq; k < 2 * wq + 2; k++) fftb_B0(xb + 12 * k * 32, StExchK{xb + 12 * k * 32, sc});
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // the previous pass's tensor store must have read its exchange buffer before the NEXT pass's stage A writes
                // it again, i.e. before anyone leaves the barrier below; waiting here (not before stage B, where the
                // store had only the length of stage A to drain) keeps the issuing warp off the critical path
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                m2_sync(P3_GRP0 + 2 * hemi + 1, 128);
                if (issuer) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(xbuf);
                    asm volatile(
                        "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tmap),
                        "r"(sa), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(ebase)
                        : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512, 1) k_spec2grid_mma3(const Ctx c, const InvDesc *__restrict__ descs, int nwork,
                                                           const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double p3_sm[];
    double *exch = p3_sm, *slots = p3_sm + 4 * M2_XH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    switch (warp) {
        case 0: s2g3_L<0>(c, descs, nwork, slots, lane); break;
        case 1: s2g3_L<1>(c, descs, nwork, slots, lane); break;
        case 2: s2g3_L<2>(c, descs, nwork, slots, lane); break;
        case 3: s2g3_L<3>(c, descs, nwork, slots, lane); break;
        case 4: s2g3_L<4>(c, descs, nwork, slots, lane); break;
        case 5: s2g3_L<5>(c, descs, nwork, slots, lane); break;
        case 6: s2g3_L<6>(c, descs, nwork, slots, lane); break;
        case 7: s2g3_L<7>(c, descs, nwork, slots, lane); break;
        default: s2g3_F(c, descs, nwork, slots, exch, &tmap, warp - 8, lane); break;
    }
}

void launch_spec2grid_mma3(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_spec2grid_mma3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_spec2grid_mma3 needs %zu bytes of shared memory per CTA (sm_100a)\n", P3_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma3<<<nwork < sms ? nwork : sms, 512, P3_SMEM, s>>>(c, d, nwork, s2g2_tensor_map(c));
}

}  // namespace spdy
