// EXPERIMENT (round 2, NOT part of the library build): measured slower than the default kernel and kept for the record.
// Result at 512 members, 73 fields (parity-green, tests/test_transforms_gpu.py): 0.589 ms against 0.505 ms for
// k_grid2spec_mma2.  Timing-only builds of THIS kernel: no stage B 0.410, no DMMAs 0.372, no stage A 0.575, nothing 0.363.
// The skeleton is faster than the old one (0.363 vs 0.430: deeper ring), but each Legendre warp now runs ~320 latency-bound
// instructions per quad in sequence (2 x 16 LDS, 2 x 102 FP64 butterflies, 32 DMMAs) and takes ~5,900 cycles for them
// (ncu: FP64 pipe 20 % + DMMA 13 % busy, i.e. not throughput; 7 + 4 + 1 = 12 warps instead of 16 cover less latency);
// sharing one code body between the five big items (instruction cache) gained 5 %, dropping the A-fragment loads nothing.
// To build it: copy next to fused_mma2.cu, include it from spdy_all.cu and call launch_grid2spec_mma3.
// speedy-b200: third-generation fused grid -> spectral transform: no Fourier slots -- the Legendre warps run FFT stage B.
//
// Reference semantics: fourier.f90:90-123 (+ fftpack.f90:136-202), legendre.f90:170-221, grid-point products of
// tendencies.f90:238-268 applied while loading.
//
// Measured on k_grid2spec_mma2 (fused_mma2.cu; 0.505 ms for the step's 73 fields at 512 members), timing-only builds:
//   without stage B 0.473, without the DMMAs 0.449, without both 0.430 -- and that skeleton (tensor loads, stage A, barriers)
//   goes 0.68 / 0.43 / 0.37 ms with a ring of 2 / 3 / 4 passes: it is bound by the bytes in flight, the full kernel by its
//   FFT warps (they issue 14 % of their cycles) while the Legendre warps wait 60 % of the time.
// The seven stage-B items of the forward FFT each produce the Fourier coefficients of ONE set of wavenumbers
//   item q: m in {q, 12-q, 12+q, 24-q, 24+q} (q = 1..5), {0, 12, 24} (q = 0), {6, 18, 30} (q = 6)      (tools/gen_fft96.py)
// and one line per lane -- lane = (latitude k of a quad, member n) -- is exactly the B fragment of the FP64 MMA (k = lane % 4,
// n = lane / 4).  So here Legendre warp q runs stage-B item q on its own line of BOTH hemispheres, folds them (E = N + S,
// O = N - S, legendre.f90:196-203) and feeds the quadrature DMMAs of those wavenumbers straight from registers:
//   * no Fourier slots (64 KB), no 62 STS + 62 LDS per line, no FULL / EMPTY hand-over per quad;
//   * the FFT warps are left with stage A only (eight 12-point items per pass, in place in the operand box);
//   * the freed shared memory deepens the ring: 6 entries for the first operand / stage-A results (three quads: one being
//     read by the Legendre warps, one in stage A, one in flight) + 3 for the second operand (dead after stage A).
// 384 threads, one CTA per SM: warps 0-6 Legendre (one stage-B item each), warp 7 requests the tensor loads, warps 8-11 FFT
// stage A (two items per pass each); setmaxnreg: 208 registers for warps 0-7, 88 for warps 8-11.
#include "fused_common.cuh"

namespace spdy {

constexpr int G3_NX = 6, G3_NY = 3, G3_NT = 384;
constexpr int G3_NL = 7 * 32, G3_NF = 4 * 32;  // Legendre / FFT threads
constexpr size_t G3_SMEM = (size_t)(G3_NX + G3_NY) * M2_XH * sizeof(double) + 64;
static_assert(G3_SMEM <= 232448, "shared memory per CTA on sm_100a");
enum { G3_STAGED0 = 1, G3_FREE0 = 4, G3_YFREE0 = 7 };  // three named barriers each

// stage-B outputs of one line stay in registers: v[r] = Fourier row r (re 2m, im 2m + 1), scaled (fourier.f90:113)
struct StReg {
    double *v;
    double sc;
    __device__ __forceinline__ void operator()(int r, double x) const { v[r] = x * sc; }
};
template <int Q> __device__ __forceinline__ void g3_stage_b(const double *__restrict__ s, const StReg st) {
    if (Q == 0) fftf_B0(s, st);
    else if (Q == 1) fftf_B1(s, st);
    else if (Q == 2) fftf_B2(s, st);
    else if (Q == 3) fftf_B3(s, st);
    else if (Q == 4) fftf_B4(s, st);
    else if (Q == 5) fftf_B5(s, st);
    else fftf_B6(s, st);
}
// quadrature tiles of wavenumber M from the folded values of this lane's latitude pair (cf. md2_mma, fused_mma2.cu)
template <int M>
__device__ __forceinline__ void md3_mma(MdC2<M> &c, const double *__restrict__ Aq, const double *v0, const double *v1) {
    constexpr int NE = MD2_NE(M), NT = MdC2<M>::NT;
    double a[NT];
#pragma unroll
#ifdef EXP3_NOALOAD
    for (int i = 0; i < NT; i++) a[i] = (double)(i + (int)(size_t)Aq);
#else
    for (int i = 0; i < NT; i++) a[i] = __ldg(Aq + (size_t)(MD2_TOFF(M) + i) * 32);
#endif
    const double er = v0[2 * M] + v1[2 * M], orr = v0[2 * M] - v1[2 * M];
    const double ei = v0[2 * M + 1] + v1[2 * M + 1], oi = v0[2 * M + 1] - v1[2 * M + 1];
#pragma unroll
    for (int i = 0; i < NT; i++) {
        dmma884(c.cr[i][0], c.cr[i][1], a[i], i < NE ? er : orr);
        dmma884(c.ci[i][0], c.ci[i][1], a[i], i < NE ? ei : oi);
    }
}

// Legendre warp of stage-B item Q
template <int Q>
__device__ __forceinline__ void g2s3_L(const Ctx &c, const FwdDesc *__restrict__ descs, const FwdOut *__restrict__ outs,
                                       const int nwork, const double *xs, const int lane, const int sparse) {
    constexpr bool SMALL = (Q == 0 || Q == 6);
    constexpr int MA = Q, MB = SMALL ? Q + 12 : 12 - Q, MC = SMALL ? Q + 24 : 12 + Q, MD = SMALL ? 0 : 24 - Q,
                  ME = SMALL ? 0 : 24 + Q;
    constexpr int NTW = MdC2<MA>::NT + MdC2<MB>::NT + MdC2<MC>::NT + (SMALL ? 0 : MdC2<MD>::NT + MdC2<ME>::NT);
    const int kk = lane & 3, col = lane >> 2;  // B fragment: latitude kk of the quad, member col
    // hemisphere 1 (pass 2g + 1) holds latitude 4jq + kk in line kk; its mirror il-1-4jq-kk is line 3 - kk of hemisphere 0
    const int l0 = (3 - kk) * 8 + col, l1 = kk * 8 + col;
    const double *pq = c.G->pq_dir2 + lane;
    // L1 prefetch of the next quad's A fragments (256 B per tile: two lines): lane -> (tile lane / 2 of this warp, half)
    int ptile = -1;
    {
        const int j = lane >> 1;
        int o = 0;
        auto pick = [&](int toff, int nt) {
            if (j >= o && j < o + nt) ptile = toff + (j - o);
            o += nt;
        };
        pick(MD2_TOFF(MA), MdC2<MA>::NT), pick(MD2_TOFF(MB), MdC2<MB>::NT), pick(MD2_TOFF(MC), MdC2<MC>::NT);
        if (!SMALL) pick(MD2_TOFF(MD), MdC2<MD>::NT), pick(MD2_TOFF(ME), MdC2<ME>::NT);
        static_assert(NTW <= 16, "one prefetch per lane");
    }
    const double *ppf = c.G->pq_dir2 + (size_t)(ptile < 0 ? 0 : ptile) * 32 + (lane & 1) * 16;
    const double sc = c_T.fc[3];
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        MdC2<MA> c0;
        MdC2<MB> c1;
        MdC2<MC> c2;
        MdC2<MD> c3;
        MdC2<ME> c4;
        md2_zero(c0), md2_zero(c1), md2_zero(c2);
        if (!SMALL) md2_zero(c3), md2_zero(c4);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int qb = g % 3;
            const double *Aq = pq + (size_t)jq * (PD2_TTOT * 32);
            // this quad's A fragments: on their way into L1 while the warp waits for the quad and runs stage B (one quad of
            // the table, 24 KB for the seven warps, is what fits next to 216 KB of shared memory)
            if (ptile >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(ppf + (size_t)jq * (PD2_TTOT * 32)));
            m2_sync(G3_STAGED0 + qb, G3_NL + G3_NF);
            double v0[M2], v1[M2];
#ifdef EXP3_NOB
            for (int i = 0; i < M2; i++) v0[i] = xs[(size_t)(2 * qb) * M2_XH + l0], v1[i] = xs[(size_t)(2 * qb + 1) * M2_XH + l1];
#else
            g3_stage_b<Q>(xs + (size_t)(2 * qb) * M2_XH + l0, StReg{v0, sc});
            g3_stage_b<Q>(xs + (size_t)(2 * qb + 1) * M2_XH + l1, StReg{v1, sc});
#endif
            m2_arrive(G3_FREE0 + qb, G3_NL + 32);  // both entries of the quad are read: the loader may overwrite them
            if (Q == 0) v0[1] = 0.0, v1[1] = 0.0;  // fourier.f90:117: Im of m = 0
#ifdef EXP3_NOMMA
            c0.cr[0][0] += v0[2 * MA] + v1[2 * MA + 1] + v0[2 * MB] + v1[2 * MB + 1] + v0[2 * MC] + v1[2 * MC + 1];
#else
            md3_mma(c0, Aq, v0, v1), md3_mma(c1, Aq, v0, v1), md3_mma(c2, Aq, v0, v1);
            if (!SMALL) md3_mma(c3, Aq, v0, v1), md3_mma(c4, Aq, v0, v1);
#endif
        }
        double *Xl = refp(c, t, outs[descs[f].fidx].dst, 0) + (size_t)(M2 * 2 * col) * TILE + MQ_NM * grp + 2 * kk;
        md2_store(c0, Xl, col, sparse), md2_store(c1, Xl, col, sparse), md2_store(c2, Xl, col, sparse);
        if (!SMALL) md2_store(c3, Xl, col, sparse), md2_store(c4, Xl, col, sparse);
    }
}

// warp 7: one thread requests the operand boxes of pass p = 2g + hemisphere into entry p % 6 (and p % 3 for the second
// operand) as soon as the Legendre warps have read quad g - 3 and the FFT warps have finished stage A of pass p - 3
__device__ __forceinline__ void g2s3_load(const Ctx &c, const FwdDesc *__restrict__ descs, const int lmode, const int nwork,
                                          double *xs, double *ys, const unsigned mbar0, const CUtensorMap *tmap, const int lane) {
    int p = 0, g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const FwdDesc d = descs[f];
        const bool two = ((lmode == FM_ALL) ? d.mode : lmode) >= FM_KE;
        const int ea = (int)((long long)t * c.scr_elems + (d.a & ~REF_SCR));
        const int eb = (int)((long long)t * c.scr_elems + (d.b & ~REF_SCR));
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
#pragma unroll 1
            for (int hemi = 0; hemi < 2; hemi++, p++) {
                if (hemi == 0 && g >= 3) m2_sync(G3_FREE0 + g % 3, G3_NL + 32);
                if (p >= G3_NY) m2_sync(G3_YFREE0 + p % G3_NY, G3_NF + 32);
                if (lane == 0) {
                    const int lat0 = hemi ? 4 * jq : IL - 4 - 4 * jq;
                    const unsigned mbar = mbar0 + 8 * (p % G3_NX);
                    const unsigned dx = (unsigned)__cvta_generic_to_shared(xs + (size_t)(p % G3_NX) * M2_XH);
                    // the entries were last accessed through the generic proxy (stage A in place, stage-B reads)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((two ? 2 : 1) * M2_XH * 8)
                                 : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dx),
                        "l"(tmap), "r"(mbar), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(ea)
                        : "memory");
                    if (two) {
                        const unsigned dy = (unsigned)__cvta_generic_to_shared(ys + (size_t)(p % G3_NY) * M2_XH);
                        asm volatile(
                            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dy),
                            "l"(tmap), "r"(mbar), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(eb)
                            : "memory");
                    }
                }
                __syncwarp();
            }
        }
    }
}

// FFT warp fw of 4: stage-A items 2 fw and 2 fw + 1 of every pass, in place on rows 12 item .. 12 item + 11 of the first
// operand's entry (lane = line: latitude lane / 8 of the box, member lane % 8)
__device__ __forceinline__ void g2s3_F(const Ctx &c, const FwdDesc *__restrict__ descs, const int lmode, const int nwork,
                                       double *xs, const double *ys, const unsigned mbar0, const int fw, const int lane) {
    const int jl = lane >> 3;
    int p = 0, g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int f = (w >> 2) / c.ntiles;
        const FwdDesc d = descs[f];
        const int mode = (lmode == FM_ALL) ? d.mode : lmode;
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
#pragma unroll 1
            for (int hemi = 0; hemi < 2; hemi++, p++) {
                const int lat = (hemi ? 4 * jq : IL - 4 - 4 * jq) + jl;
                double *ba = xs + (size_t)(p % G3_NX) * M2_XH + lane;
                const double *bb = ys + (size_t)(p % G3_NY) * M2_XH + lane;
                const double sc = (d.kcos == 3) ? c_T.cosgr2[lat] : c_T.cosgr[lat];
                g2_mbar_wait(mbar0 + 8 * (p % G3_NX), (p / G3_NX) & 1);
#ifdef EXP3_NOA
                if (0)
#endif
#pragma unroll 1
                for (int it = 2 * fw; it < 2 * fw + 2; it++) {
                    double *sa = ba + 12 * it * 32;
                    const double *sb = bb + 12 * it * 32;
                    if (mode == FM_PLAIN) fftf_A0(LdBox<FM_PLAIN>{sa, sa, d.k0, sc}, sa);
                    else if (mode == FM_COS) fftf_A0(LdBox<FM_COS>{sa, sa, d.k0, sc}, sa);
                    else if (mode == FM_KE) fftf_A0(LdBox<FM_KE>{sa, sb, d.k0, sc}, sa);
                    else if (mode == FM_FLUXT) fftf_A0(LdBox<FM_FLUXT>{sa, sb, d.k0, sc}, sa);
                    else fftf_A0(LdBox<FM_FLUX>{sa, sb, d.k0, sc}, sa);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the next access to both entries is a tensor load
                m2_arrive(G3_YFREE0 + p % G3_NY, G3_NF + 32);  // the second operand's entry is dead
                if (hemi == 1) m2_arrive(G3_STAGED0 + g % 3, G3_NL + G3_NF);  // both hemispheres of quad g are staged
            }
        }
    }
}

__global__ void __launch_bounds__(G3_NT, 1) k_grid2spec_mma3(const Ctx c, const FwdDesc *__restrict__ descs,
                                                            const FwdOut *__restrict__ outs, int nwork,
                                                            const __grid_constant__ CUtensorMap tmap, int sparse, int lmode) {
    extern __shared__ __align__(128) double g3_sm[];
    double *xs = g3_sm, *ys = g3_sm + (size_t)G3_NX * M2_XH;
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(ys + (size_t)G3_NY * M2_XH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < G3_NX) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8 * threadIdx.x));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp >= 8) {
        reg_dec<88>();
        g2s3_F(c, descs, lmode, nwork, xs, ys, mbar0, warp - 8, lane);
    } else {
        reg_inc<208>();
        switch (warp) {  // the two light items (0, 6) share schedulers with heavy ones
            case 0: g2s3_L<1>(c, descs, outs, nwork, xs, lane, sparse); break;
#ifndef EXP3_SAME
            case 1: g2s3_L<2>(c, descs, outs, nwork, xs, lane, sparse); break;
            case 2: g2s3_L<3>(c, descs, outs, nwork, xs, lane, sparse); break;
            case 3: g2s3_L<4>(c, descs, outs, nwork, xs, lane, sparse); break;
#else
            case 1: case 2: case 3: case 6: g2s3_L<1>(c, descs, outs, nwork, xs, lane, sparse); break;
#endif
            case 4: g2s3_L<0>(c, descs, outs, nwork, xs, lane, sparse); break;
            case 5: g2s3_L<6>(c, descs, outs, nwork, xs, lane, sparse); break;
#ifndef EXP3_SAME
            case 6: g2s3_L<5>(c, descs, outs, nwork, xs, lane, sparse); break;
#endif
            default: g2s3_load(c, descs, lmode, nwork, xs, ys, mbar0, &tmap, lane); break;
        }
    }
}

// all operand fields must live in the scratch arena (true for the model step's lists and the batch workspace)
void launch_grid2spec_mma3(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, const FwdOut *o, int nf, int sparse) {
    if (!nf) return;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_grid2spec_mma3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G3_SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_grid2spec_mma3 needs %zu bytes of shared memory per CTA (sm_100a)\n", G3_SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_grid2spec_mma3<<<nwork < sms ? nwork : sms, G3_NT, G3_SMEM, s>>>(c, d, o, nwork, s2g2_tensor_map(c), sparse, mode);
}

}  // namespace spdy
