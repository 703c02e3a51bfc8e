// speedy-b200: second-generation fused spectral -> grid transform (SPDY_FUSED=5).
//
// Reference semantics: legendre.f90:130-168 (inverse Legendre), fourier.f90:63-88 + fftpack.f90:69-134 (inverse FFT).
//
// Same decomposition as k_spec2grid_mma (fused_mma.cu): work item = (field, tile, 8-member group), Legendre
// contraction as mma.sync.m8n8k4.f64 per (wavenumber, latitude quad), the Fourier rows of a quad (8 latitude rows x
// 62 x 8 members) handed over in a shared-memory slot.  What changed is the occupancy of both sides, because ncu
// and the timing experiments (profiles/README.md) showed the first kernel bound by latency with ONE Legendre warp
// and ONE whole-line FFT warp (230 registers) per scheduler:
//   * 8 "L" warps, each owning 3-4 wavenumbers (15-20 k-slices): B fragments of a work item in 72-80 registers;
//   * 8 "F" warps running the TWO-STAGE FFT of the separate kernel (fft96_gen.cuh items, <= 16 points per thread)
//     with the exchange through a second shared-memory buffer: 4 warps per hemisphere (= 32 lines: 4 latitude rows
//     x 8 members), stage A items split 2+2+1+2, the eight identical 12-point stage-B items two per warp.
//   Everything fits 128 registers per thread, so the CTA has 16 warps (4 per scheduler) instead of 8.
// Shared memory: NS slots of 31.5 KB + 2 hemispheres x 2 (double buffer) x 24 KB exchange = 224 KB with NS = 4.
// Hand-over: named barriers FULL[slot] (L arrive, F sync), EMPTY[slot] (F arrive, L sync), one 128-thread barrier
// per hemisphere group between stage A and stage B (the double-buffered exchange needs no second one).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace spdy {

#ifndef M2_NS
#define M2_NS 2
#endif
constexpr int M2_NSLOT = M2_NS;
constexpr int M2_XH = IX * 32;                                   // doubles per exchange buffer (32 lines)
constexpr size_t M2_SMEM = ((size_t)M2_NSLOT * MQ_SLOT + 4 * M2_XH) * sizeof(double);
enum { M2_FULL0 = 1, M2_EMPTY0 = 1 + M2_NSLOT, M2_GRP0 = 1 + 2 * M2_NSLOT };  // + 4 group barriers
static_assert(M2_GRP0 + 3 <= 15, "named barriers");
static_assert(M2_SMEM <= 232448, "shared memory per CTA on sm_100a");

__device__ __forceinline__ void m2_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#ifndef M2_EXP
#define M2_EXP 0
#endif
__device__ __forceinline__ void m2_arrive(int id, int n) {
#ifdef MQ_FENCE
    __threadfence_block();
#endif
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

template <int M> __device__ __forceinline__ void mq_prefetch_b(const double *__restrict__ Xl) {
#pragma unroll
    for (int s = 0; s < MQ_KS(M); s++) {
        const double *x = Xl + (size_t)((2 * M) + M2 * 4 * s) * TILE;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(x));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + TILE));
    }
}

// L warp LW of 8: wavenumber pairs (LW, 30-LW) and (LW+8, 22-LW); LW = 7: (7, 23) and 15 alone.
template <int LW>
__device__ __forceinline__ void s2g2_L(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork, double *slots,
                                       const int lane) {
    constexpr int M3 = (LW != 7) ? 22 - LW : 15;
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_inv + (lane & 15);
    const long long sgn = (lane >= 16 && (lane & 1)) ? (long long)0x8000000000000000ull : 0ll;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const double *Xl = refp(c, t, descs[f].src, 0) + (size_t)(M2 * kk) * TILE + MQ_NM * grp + col;
        MqB<LW> b0;
        MqB<30 - LW> b1;
        MqB<LW + 8> b2;
        MqB<M3> b3;
        mq_load_b(b0, Xl, kk), mq_load_b(b1, Xl, kk), mq_load_b(b2, Xl, kk);
        if (LW != 7) mq_load_b(b3, Xl, kk);
        if (w + (int)gridDim.x < nwork) {  // the coefficients of this warp's next work item: pull them into L2 now
            const int wn = w + gridDim.x, tn = (wn >> 2) % c.ntiles, fn = (wn >> 2) / c.ntiles;
            const double *Xn = refp(c, tn, descs[fn].src, 0) + (size_t)(M2 * kk) * TILE + MQ_NM * (wn & 3) + col;
            mq_prefetch_b<LW>(Xn), mq_prefetch_b<30 - LW>(Xn), mq_prefetch_b<LW + 8>(Xn);
            if (LW != 7) mq_prefetch_b<M3>(Xn);
        }
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % M2_NSLOT;
            double *Sl = slots + sl * MQ_SLOT + col * MQ_RS + 2 * kk;
            const double *Aq = pq + (size_t)jq * MQ_KTOT * 16;
            MqA<LW> a0;
            MqA<30 - LW> a1;
            MqA<LW + 8> a2;
            MqA<M3> a3;
            if (!(M2_EXP & 1)) {
                mq_load_a(a0, Aq, sgn), mq_load_a(a1, Aq, sgn), mq_load_a(a2, Aq, sgn);
                if (LW != 7) mq_load_a(a3, Aq, sgn);
            }
            if (g >= M2_NSLOT) m2_sync(M2_EMPTY0 + sl, 512);
            if (!(M2_EXP & 1)) {
                mq_mma_store2(a0, b0, a1, b1, Sl);
                if (LW != 7) mq_mma_store2(a2, b2, a3, b3, Sl);
                else mq_mma_store(a2, b2, Sl);
            }
            m2_arrive(M2_FULL0 + sl, 512);
        }
    }
}

// stage-B outputs go back IN PLACE into the exchange rows the thread has just read (item k: rows 12k..12k+11 hold
// grid points k + 8q, q = 0..11, afterwards), scaled for kcos = 2 (fourier.f90:88-92)
struct StExchK {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int i, double v) const { p[(i >> 3) * 32] = v * sc; }
};

// F warp fw of 8: hemisphere fw >> 2, item share fw & 3; lane = (jl, member).  Thread jl of hemisphere 1 owns slot
// row 4 + jl (latitude 4jq + jl), of hemisphere 0 slot row 3 - jl (latitude il-4-4jq + jl, legendre.f90:163-167), so the
// four latitude rows of a hemisphere group ascend with jl and the 24 KB exchange buffer [row 12k+q][jl][member] is
// exactly one TMA box [8 members][4 latitudes][12 q][8 k] of the scratch arena seen as a 5-D tensor (see the launcher).
// The grid values leave the SM by ONE cp.async.bulk.tensor store per hemisphere and quad, issued by one thread: no F
// warp ever waits in the LSU queue (with per-thread STG the store drain was 0.15 ms of this kernel).
__device__ __forceinline__ void s2g2_F(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                       const double *slots, double *exch, const CUtensorMap *tmap, const int fw,
                                       const int lane) {
    const int hemi = fw >> 2, wq = fw & 3, jl = lane >> 3, mem = lane & 7, row = hemi ? 4 + jl : 3 - jl;
    const bool issuer = (wq == 0 && lane == 0);
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const InvDesc d = descs[f];
        const int ebase = (int)((long long)t * c.scr_elems + d.dst);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % M2_NSLOT;
            const int lat0 = hemi ? 4 * jq : IL - 4 - 4 * jq, lat = lat0 + jl;
            const LdSlot ld{slots + sl * MQ_SLOT + row * MQ_RS + mem};
            double *xbuf = exch + (size_t)(2 * hemi + (g & 1)) * M2_XH, *xb = xbuf + lane;
            m2_sync(M2_FULL0 + sl, 512);
            if (M2_EXP & 2) {
                m2_arrive(M2_EMPTY0 + sl, 512);
                continue;
            }
            if (wq == 0) {
                fftb_A1(ld, xb), fftb_A0(ld, xb);
            } else if (wq == 1) {
                fftb_A2(ld, xb), fftb_A6(ld, xb);
            } else if (wq == 2) {
                fftb_A3(ld, xb), fftb_A4(ld, xb);
            } else {
                fftb_A5(ld, xb);
            }
            m2_arrive(M2_EMPTY0 + sl, 512);
            // the bulk store of the previous quad (other exchange buffer) must have read its buffer before any warp of
            // the group starts stage A of the next quad: the issuer checks before it joins this barrier
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            m2_sync(M2_GRP0 + 2 * hemi, 128);
            const double sc = d.kcos == 1 ? 1.0 : c_T.cosgr[lat];
#pragma unroll 1
            for (int k = 2 * wq; k < 2 * wq + 2; k++) fftb_B0(xb + 12 * k * 32, StExchK{xb + 12 * k * 32, sc});
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            m2_sync(M2_GRP0 + 2 * hemi + 1, 128);
            if (issuer && !(M2_EXP & 4)) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(xbuf);
                asm volatile(
                    "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tmap),
                    "r"(sa), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(ebase)
                    : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512, 1) k_spec2grid_mma2(const Ctx c, const InvDesc *__restrict__ descs, int nwork,
                                                           const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double m2_sm[];
    double *exch = m2_sm + M2_NSLOT * MQ_SLOT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    switch (warp) {
        case 0: s2g2_L<0>(c, descs, nwork, m2_sm, lane); break;
        case 1: s2g2_L<1>(c, descs, nwork, m2_sm, lane); break;
        case 2: s2g2_L<2>(c, descs, nwork, m2_sm, lane); break;
        case 3: s2g2_L<3>(c, descs, nwork, m2_sm, lane); break;
        case 4: s2g2_L<4>(c, descs, nwork, m2_sm, lane); break;
        case 5: s2g2_L<5>(c, descs, nwork, m2_sm, lane); break;
        case 6: s2g2_L<6>(c, descs, nwork, m2_sm, lane); break;
        case 7: s2g2_L<7>(c, descs, nwork, m2_sm, lane); break;
        default: s2g2_F(c, descs, nwork, m2_sm, exch, &tmap, warp - 8, lane); break;
    }
}

// The scratch arena as a 5-D FP64 tensor for the TMA stores: element (lane, lat, q, k, e) lives at
//   scr + 8 * lane + 256 * (e + 96 * lat + 8 * q + k)   bytes,
// i.e. grid point i = k + 8q of latitude row lat of the (96,48) field that starts at tile-relative element offset e
// (lane stride 8 B, point stride 256 B: the 32-member tile layout of spdy.cuh).  Box = [8][4][12][8][1].
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static const CUtensorMap &s2g2_tensor_map(const Ctx &c) {
    static EncodeTiledFn encode = nullptr;
    static CUtensorMap map;
    static const void *k_scr = nullptr;
    static long long k_elems = -1;
    static int k_tiles = -1;
    if (!encode) {
        cudaDriverEntryPointQueryResult qr;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
            fprintf(stderr, "speedy_b200: cuTensorMapEncodeTiled is not available in this driver\n");
            abort();
        }
        encode = (EncodeTiledFn)fn;
    }
    if (k_scr != c.scr || k_elems != c.scr_elems || k_tiles != c.ntiles) {
        const cuuint64_t dims[5] = {(cuuint64_t)TILE, (cuuint64_t)IL, 12, 8, (cuuint64_t)c.ntiles * (cuuint64_t)c.scr_elems};
        const cuuint64_t strides[4] = {(cuuint64_t)IX * TILE * 8, 8ull * TILE * 8, (cuuint64_t)TILE * 8, (cuuint64_t)TILE * 8};
        const cuuint32_t box[5] = {MQ_NM, 4, 12, 8, 1}, estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, c.scr, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fprintf(stderr, "speedy_b200: cuTensorMapEncodeTiled failed (%d)\n", (int)r);
            abort();
        }
        k_scr = c.scr, k_elems = c.scr_elems, k_tiles = c.ntiles;
    }
    return map;
}

void launch_spec2grid_mma2(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_spec2grid_mma2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M2_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_spec2grid_mma2 needs %zu bytes of shared memory per CTA (sm_100a)\n", M2_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma2<<<nwork < sms ? nwork : sms, 512, M2_SMEM, s>>>(c, d, nwork, s2g2_tensor_map(c));
}

}  // namespace spdy
