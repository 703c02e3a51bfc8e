// EXPERIMENT (round 2, NOT part of the library build): measured slower than the default kernel and kept for the record.
// Result at 512 members, 77 fields (parity-green): 0.647 ms against 0.444 ms for k_spec2grid_mma4 (fused_mma4.cu): one
// item per warp and stage leaves the 24..96-flop stage-A items unbalanced behind a 256-thread barrier, and 136 registers
// for two lines per thread spill (192 bytes of stack).
// speedy-b200: fifth-generation fused spectral -> grid transform: two FFT lines per thread.
//
// Reference semantics: legendre.f90:130-168 (inverse Legendre), fourier.f90:63-88 + fftpack.f90:69-134 (inverse FFT).
//
// ncu of k_spec2grid_mma4 (12 FFT warps): the FFT warps were stalled by the shared-memory instruction queue (mio_throttle 2.5
// per issued instruction, LSU wavefronts 58 %), not by their critical path: one LDS.64 / STS.64 per 3 FP64 instructions.
// Here a thread owns TWO lines -- the same latitude, two neighbouring members -- so that every shared-memory access of
// an FFT warp is 128 bits wide (half the LDS / STS instructions per line) and every butterfly chain has an independent
// twin (the chains are latency-bound: 8.5 cycles per dependent FP64 instruction).  A pass is a whole hemisphere octet
// (8 latitudes x 8 members = 64 lines, 48 KB exchange buffer, double-buffered) worked on by ALL eight FFT warps: stage A
// one item per warp (A1..A5, A0, A6; one warp idle), one 256-thread barrier, stage B exactly one of the eight 12-point
// items per warp, in place, then ONE TMA tensor store of the 48 KB box [8 members][8 latitudes][12][8].
// 4 Legendre (DMMA) warps at 232 registers + 8 FFT warps at 136 (setmaxnreg), 384 threads, one CTA per SM.
#include "fused_common.cuh"

namespace spdy {

// Two FFT lines per thread (the same latitude, two neighbouring members): the generated butterfly items (fft96_gen.cuh) are
// templates on the value type, and with D2 every shared-memory access of an FFT warp is 128 bits wide -- half the LDS /
// STS instructions per line and two independent dependency chains per thread.
struct __align__(16) D2 {
    double x, y;
};
__device__ __forceinline__ D2 operator+(const D2 a, const D2 b) { return D2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ D2 operator-(const D2 a, const D2 b) { return D2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ D2 operator-(const D2 a) { return D2{-a.x, -a.y}; }
__device__ __forceinline__ D2 operator*(const double s, const D2 a) { return D2{s * a.x, s * a.y}; }
__device__ __forceinline__ D2 operator*(const D2 a, const double s) { return D2{a.x * s, a.y * s}; }
struct LdSlot2 {  // stage-A loader for a pair of members: Fourier row r of a slot
    const double *p;
    __device__ __forceinline__ D2 operator()(int r) const { return *reinterpret_cast<const D2 *>(p + r * MQ_NM); }
};
struct StExch2 {  // in-place stage-B store for a pair of lines (see StExchK), optional 1/cos(lat) factor
    D2 *p;
    double sc;
    bool scale;
    __device__ __forceinline__ void operator()(int i, D2 v) const { p[(i >> 3) * 32] = scale ? v * sc : v; }
};


constexpr int P5_XH = IX * 64;  // doubles per exchange buffer (64 lines)
constexpr size_t P5_SMEM = ((size_t)2 * P3_SLOT + 2 * P5_XH) * sizeof(double);
static_assert(P5_SMEM <= 232448, "shared memory per CTA on sm_100a");
enum { P5_FULL0 = 1, P5_EMPTY0 = 3, P5_GRP0 = 5 };  // + 2 group barriers
constexpr int P5_NT = 384;

template <int LW>
__device__ __forceinline__ void s2g5_L(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork, double *slots,
                                       const int lane) {
    constexpr int M7 = (LW != 3) ? 18 - LW : 15;
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_inv2 + lane;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const double *Xl = refp(c, t, descs[f].src, 0) + MQ_NM * grp + col;
        P3B<LW> b0;
        P3B<30 - LW> b1;
        P3B<LW + 8> b2;
        P3B<22 - LW> b3;
        P3B<LW + 4> b4;
        P3B<26 - LW> b5;
        P3B<LW + 12> b6;
        P3B<M7> b7;
        p3_load_b(b0, Xl, kk), p3_load_b(b1, Xl, kk), p3_load_b(b2, Xl, kk), p3_load_b(b3, Xl, kk);
        p3_load_b(b4, Xl, kk), p3_load_b(b5, Xl, kk), p3_load_b(b6, Xl, kk);
        if (LW != 3) p3_load_b(b7, Xl, kk);
        if (w + (int)gridDim.x < nwork) {  // the coefficients of this warp's next work item: pull them into L2 now
            const int wn = w + gridDim.x, tn = (wn >> 2) % c.ntiles, fn = (wn >> 2) / c.ntiles;
            const double *Xn = refp(c, tn, descs[fn].src, 0) + MQ_NM * (wn & 3) + col;
            p3_prefetch_b<LW>(Xn, kk), p3_prefetch_b<30 - LW>(Xn, kk), p3_prefetch_b<LW + 8>(Xn, kk), p3_prefetch_b<22 - LW>(Xn, kk);
            p3_prefetch_b<LW + 4>(Xn, kk), p3_prefetch_b<26 - LW>(Xn, kk), p3_prefetch_b<LW + 12>(Xn, kk);
            if (LW != 3) p3_prefetch_b<M7>(Xn, kk);
        }
#pragma unroll 1
        for (int jo = 0; jo < IY / 8; jo++, g++) {
            const int sl = g & 1;
            double *Sl = slots + sl * P3_SLOT + col * MQ_RS + 2 * kk;
            const double *Aq = pq + (size_t)jo * PQ2_KTOT * 32;
            P3A<LW> a0;
            P3A<30 - LW> a1;
            P3A<LW + 8> a2;
            P3A<22 - LW> a3;
            P3A<LW + 4> a4;
            P3A<26 - LW> a5;
            P3A<LW + 12> a6;
            P3A<M7> a7;
            p3_load_a(a0, Aq);
            if (g >= 2) m2_sync(P5_EMPTY0 + sl, P5_NT);
            p3_load_a(a1, Aq);
            p3_mma_store(b0, a0, Sl);
            p3_load_a(a2, Aq);
            p3_mma_store(b1, a1, Sl);
            p3_load_a(a3, Aq);
            p3_mma_store(b2, a2, Sl);
            p3_load_a(a4, Aq);
            p3_mma_store(b3, a3, Sl);
            p3_load_a(a5, Aq);
            p3_mma_store(b4, a4, Sl);
            p3_load_a(a6, Aq);
            p3_mma_store(b5, a5, Sl);
            if (LW != 3) p3_load_a(a7, Aq);
            p3_mma_store(b6, a6, Sl);
            if (LW != 3) p3_mma_store(b7, a7, Sl);
            m2_arrive(P5_FULL0 + sl, P5_NT);
        }
    }
}

// F warp fw of 8; lane = (latitude l8 = lane / 4 of the hemisphere's eight, member pair mp = lane % 4).  Hemisphere 1 rows
// 8 + l8 of the slot hold latitude 8jo + l8; hemisphere 0 is read in reverse (slot row 7 - l8 = latitude il-8-8jo + l8) so
// that the eight latitudes of a pass ascend with l8: one TMA box.
__device__ __forceinline__ void s2g5_F(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                       const double *slots, double *exch, const CUtensorMap *tmap, const int fw,
                                       const int lane) {
    const int l8 = lane >> 2, mp = lane & 3;
    const bool issuer = (fw == 7 && lane == 0);  // warp 7 has no stage-A item
    int g = 0, p = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const InvDesc d = descs[f];
        const int ebase = (int)((long long)t * c.scr_elems + d.dst);
#pragma unroll 1
        for (int jo = 0; jo < IY / 8; jo++, g++) {
            const int sl = g & 1;
            m2_sync(P5_FULL0 + sl, P5_NT);
#pragma unroll 1
            for (int hemi = 0; hemi < 2; hemi++, p++) {
                const int row = hemi ? 8 + l8 : 7 - l8;
                const int lat0 = hemi ? 8 * jo : IL - 8 - 8 * jo, lat = lat0 + l8;
                const LdSlot2 ld{slots + sl * P3_SLOT + row * MQ_RS + 2 * mp};
                double *xbuf = exch + (size_t)(p & 1) * P5_XH;
                D2 *xb = reinterpret_cast<D2 *>(xbuf) + lane;
                switch (fw) {  // warp-uniform
                    case 0: fftb_A1(ld, xb); break;
                    case 1: fftb_A2(ld, xb); break;
                    case 2: fftb_A3(ld, xb); break;
                    case 3: fftb_A4(ld, xb); break;
                    case 4: fftb_A5(ld, xb); break;
                    case 5: fftb_A0(ld, xb); break;
                    case 6: fftb_A6(ld, xb); break;
                    default: break;
                }
                if (hemi == 1) m2_arrive(P5_EMPTY0 + sl, P5_NT);  // this warp has read its share of the slot completely
                m2_sync(P5_GRP0, 256);
                fftb_B0(xb + 12 * fw * 32, StExch2{xb + 12 * fw * 32, c_T.cosgr[lat], d.kcos != 1});
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // the previous pass's tensor store must have read its exchange buffer before the NEXT pass's stage A writes
                // it again, i.e. before anyone leaves the barrier below
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                m2_sync(P5_GRP0 + 1, 256);
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

__global__ void __launch_bounds__(P5_NT, 1) k_spec2grid_mma5(const Ctx c, const InvDesc *__restrict__ descs, int nwork,
                                                             const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double p5_sm[];
    double *exch = p5_sm, *slots = p5_sm + 2 * P5_XH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 4) {  // two FFT warpgroups give registers back ...
        reg_dec<136>();
        s2g5_F(c, descs, nwork, slots, exch, &tmap, warp - 4, lane);
    } else {  // ... to the Legendre warpgroup
        reg_inc<232>();
        switch (warp) {
            case 0: s2g5_L<0>(c, descs, nwork, slots, lane); break;
            case 1: s2g5_L<1>(c, descs, nwork, slots, lane); break;
            case 2: s2g5_L<2>(c, descs, nwork, slots, lane); break;
            default: s2g5_L<3>(c, descs, nwork, slots, lane); break;
        }
    }
}

void launch_spec2grid_mma5(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_spec2grid_mma5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P5_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_spec2grid_mma5 needs %zu bytes of shared memory per CTA (sm_100a)\n", P5_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma5<<<nwork < sms ? nwork : sms, P5_NT, P5_SMEM, s>>>(c, d, nwork, s2g2_tensor_map(c, 8));
}

}  // namespace spdy
