// speedy-b200: fused spectral -> grid transform, Legendre side: parity-pure DMMA over latitude octets (fragments and the
// accumulate-and-store step shared by the kernel in fused_mma4.cu).
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
// A slot is an octet (16 latitude rows x 62 x 8 members = 63 KB), two slots; the FFT side (fused_mma4.cu) takes a slot in
// two passes per hemisphere (4 latitude rows x 8 members each).  224 KB of shared memory.
// Measured (512 members, 77 fields): 0.469 ms with 8 Legendre + 8 FFT warps (round 1) against 0.496 with the N/S symmetry
// folded into M; 0.444 ms with the 4 + 12 warp split of fused_mma4.cu (round 2).
#include "fused_common.cuh"

namespace spdy {

__host__ __device__ constexpr int P3_KSE(int m) { return ((32 - m + 1) / 2 + 3) / 4; }  // k-slices of even n (n = 0, 2, ..)
__host__ __device__ constexpr int P3_KSO(int m) { return ((32 - m) / 2 + 3) / 4; }      // k-slices of odd n
__host__ __device__ constexpr int P3_KOFF(int m) { int o = 0; for (int i = 0; i < m; i++) o += P3_KSE(i) + P3_KSO(i); return o; }
static_assert(P3_KOFF(MX) == PQ2_KTOT, "GlobTables::pq_inv2 layout");
constexpr int P3_SLOT = 16 * MQ_RS;                                   // doubles per slot (octet)
constexpr size_t P3_SMEM = ((size_t)2 * P3_SLOT + 4 * M2_XH) * sizeof(double);
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

struct StExch1 {  // StExchK (fused_common.cuh) without the scaling
    double *p;
    __device__ __forceinline__ void operator()(int i, double v) const { p[(i >> 3) * 32] = v; }
};

}  // namespace spdy
