// EXPERIMENT (round 2, NOT part of the library build): measured slower than the default kernel and kept for the record.
// k_spec2grid_mma4 with SIXTEEN FFT warps (two groups of eight: one stage-B item per warp, per-pass critical path 170 flop
// units instead of 244) at 64 registers + 4 Legendre warps at 224 (640 threads).  Parity-green; 0.797 ms against 0.448 ms:
// the 16-point stage-A items do not fit 64 registers (ptxas: 320 bytes of spill stores, 328 of spill loads per thread).
// speedy-b200: fourth-generation fused spectral -> grid transform: 4 Legendre (DMMA) warps + 12 FFT warps.
//
// Reference semantics: legendre.f90:130-168 (inverse Legendre), fourier.f90:63-88 + fftpack.f90:69-134 (inverse FFT).
//
// Same work decomposition, slots, exchange buffers, parity-pure DMMA k-slices and TMA tensor stores as the round-1 kernel k_spec2grid_mma3
// (fused_mma3.cu).  What changed is the split of the register file (ncu of mma3: the FFT warps are the bottleneck -- the
// Legendre warps spend half their time waiting for an empty slot -- and their four-warp groups are badly balanced: the
// seven stage-A items cost 24..96 flops, so the 2+2+2+1 split has a critical path of 192 flop units per pass against 96
// for the heaviest single item, and the eight stage-B items leave no slack either):
//   * ONE warpgroup of 4 Legendre warps, each owning 7-8 wavenumbers (35-40 k-slices: all their B fragments in registers
//     for the whole work item), raised to 232 registers per thread with setmaxnreg;
//   * THREE warpgroups = 12 FFT warps at 88 registers, as two groups of SIX per hemisphere: stage A one big item per warp
//     (A1..A5, and A0 + A6 together on the sixth: critical path 96 instead of 192), stage B 2+2+1+1+1+1 items
//     (critical path unchanged, 148): 244 instead of 340 flop units per pass.
// 512 threads, one CTA per SM, 224 KB of shared memory as before.
#include "fused_common.cuh"

namespace spdy {

enum { P6_FULL0 = 1, P6_EMPTY0 = 3, P6_GRP0 = 5 };  // + 4 group barriers
constexpr int P6_NT = 640;


// L warp LW of 4: the wavenumbers of the old warps LW and LW + 4:
//   LW, 30-LW, LW+8, 22-LW  and  LW+4, 26-LW, LW+12, 18-LW (LW = 3: 7, 23, 15 for the second set)
template <int LW>
__device__ __forceinline__ void s2g6_L(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork, double *slots,
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
            if (g >= 2) m2_sync(P6_EMPTY0 + sl, P6_NT);
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
            m2_arrive(P6_FULL0 + sl, P6_NT);
        }
    }
}

// F warp fw of 12: hemisphere fw / 6, item share fw % 6; lane = (jl, member); two passes (halves of the hemisphere's eight
// latitudes) per octet, slot rows, exchange buffers and the TMA store as in the round-1 kernel
__device__ __forceinline__ void s2g6_F(const Ctx &c, const InvDesc *__restrict__ descs, const int nwork,
                                       const double *slots, double *exch, const CUtensorMap *tmap, const int fw,
                                       const int lane) {
    const int hemi = fw >> 3, wq = fw & 7, jl = lane >> 3, mem = lane & 7;
    const bool issuer = (wq == 7 && lane == 0);  // a warp without a stage-A item
    const int kb0 = wq, kb1 = wq + 1;  // one stage-B item per warp
    int g = 0, p = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const InvDesc d = descs[f];
        const int ebase = (int)((long long)t * c.scr_elems + d.dst);
#pragma unroll 1
        for (int jo = 0; jo < IY / 8; jo++, g++) {
            const int sl = g & 1;
            m2_sync(P6_FULL0 + sl, P6_NT);
#pragma unroll 1
            for (int half = 0; half < 2; half++, p++) {
                const int row = hemi ? 8 + 4 * half + jl : 4 * half + 3 - jl;
                const int lat0 = hemi ? 8 * jo + 4 * half : IL - 4 - 8 * jo - 4 * half, lat = lat0 + jl;
                const LdSlot ld{slots + sl * P3_SLOT + row * MQ_RS + mem};
                double *xbuf = exch + (size_t)(2 * hemi + (p & 1)) * M2_XH, *xb = xbuf + lane;
                if (wq == 0) {
                    fftb_A1(ld, xb);
                } else if (wq == 1) {
                    fftb_A2(ld, xb);
                } else if (wq == 2) {
                    fftb_A3(ld, xb);
                } else if (wq == 3) {
                    fftb_A4(ld, xb);
                } else if (wq == 4) {
                    fftb_A5(ld, xb);
                } else if (wq == 5) {
                    fftb_A0(ld, xb), fftb_A6(ld, xb);
                }
                if (half == 1) m2_arrive(P6_EMPTY0 + sl, P6_NT);  // this warp has read its share of the slot completely
                m2_sync(P6_GRP0 + 2 * hemi, 256);
                if (d.kcos == 1) {  // 59 of the 77 fields: no 1/cos(lat) factor, no multiply per grid point
#pragma unroll 1
                    for (int k = kb0; k < kb1; k++) fftb_B0(xb + 12 * k * 32, StExch1{xb + 12 * k * 32});
                } else {
                    const double sc = c_T.cosgr[lat];
#pragma unroll 1
                    for (int k = kb0; k < kb1; k++) fftb_B0(xb + 12 * k * 32, StExchK{xb + 12 * k * 32, sc});
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // the previous pass's tensor store must have read its exchange buffer before the NEXT pass's stage A writes
                // it again, i.e. before anyone leaves the barrier below
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                m2_sync(P6_GRP0 + 2 * hemi + 1, 256);
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

__global__ void __launch_bounds__(P6_NT, 1) k_spec2grid_mma6(const Ctx c, const InvDesc *__restrict__ descs, int nwork,
                                                           const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double p3_sm[];
    double *exch = p3_sm, *slots = p3_sm + 4 * M2_XH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 4) {  // three FFT warpgroups give registers back ...
        reg_dec<64>();
        s2g6_F(c, descs, nwork, slots, exch, &tmap, warp - 4, lane);
    } else {  // ... to the Legendre warpgroup
        reg_inc<224>();
        switch (warp) {
            case 0: s2g6_L<0>(c, descs, nwork, slots, lane); break;
            case 1: s2g6_L<1>(c, descs, nwork, slots, lane); break;
            case 2: s2g6_L<2>(c, descs, nwork, slots, lane); break;
            default: s2g6_L<3>(c, descs, nwork, slots, lane); break;
        }
    }
}

void launch_spec2grid_mma6(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_spec2grid_mma6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_spec2grid_mma6 needs %zu bytes of shared memory per CTA (sm_100a)\n", P3_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_spec2grid_mma6<<<nwork < sms ? nwork : sms, P6_NT, P3_SMEM, s>>>(c, d, nwork, s2g2_tensor_map(c));
}

}  // namespace spdy
