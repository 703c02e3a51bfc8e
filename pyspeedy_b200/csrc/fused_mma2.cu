// speedy-b200: fused grid -> spectral transform (FFT + Gaussian quadrature on the FP64 tensor cores, one kernel).
#include "fused_common.cuh"

namespace spdy {

// Fused grid -> spectral transform (default path).  Reference semantics: fourier.f90:90-123
// (+ fftpack.f90:136-202), legendre.f90:170-221, grid-point products of tendencies.f90:238-268 applied while loading.
//
// Mirror of k_spec2grid_mma4 (fused_mma4.cu):
//   * the grid rows of a hemisphere-quad (4 latitudes x 96 points x 8 members, 24 KB per operand field) arrive by ONE
//     cp.async.bulk.tensor load per operand through the 5-D tensor map of fused_common.cuh, two passes ahead of
//     their use (ring of three buffers, mbarrier completion): no F warp ever waits for a global load;
//   * the box arrives as [row 12k+q][latitude][member] = grid point k+8q, which is exactly the input order of the
//     eight identical 12-point stage-A items: F warp k runs item k IN PLACE on rows 12k..12k+11 (products of the two
//     operand buffers formed while reading), one 256-thread barrier, then the seven stage-B items write the 62 Fourier
//     rows of the pass's hemisphere into the slot;
//   * 8 F warps work on ONE hemisphere (32 lines) per pass, two passes per latitude quad; 8 L warps accumulate the
//     Gaussian quadrature as DMMA over the six quads of a work item (10 n-tiles x {re,im} per warp = 80 registers).
constexpr int G2_NS = 2;                       // slots
template <int MODE> struct G2Cfg {
    // FM_ALL: one launch for a mixed list (FwdDesc::mode per field), ring entries sized for two operands
    static constexpr int NOP = (MODE == FM_KE || MODE == FM_FLUXT || MODE == FM_FLUX || MODE == FM_ALL) ? 2 : 1;
    static constexpr int NB = (NOP == 2) ? 3 : 6;  // input buffer ring (passes): what fits next to the slots
    static constexpr size_t SMEM = ((size_t)G2_NS * MD_SLOT + (size_t)NB * NOP * M2_XH) * sizeof(double) + 64;
};
enum { G2_FULL0 = 1, G2_EMPTY0 = 1 + G2_NS, G2_GRP = 1 + 2 * G2_NS };

__device__ __forceinline__ void g2_mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nG2_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra G2_WAIT;\n}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}

// Parity-pure accumulator tiles: the L warp folds the two hemispheres of a B fragment (E = N + S, O = N - S,
// legendre.f90:196-203) and multiplies the even-n tiles by E and the odd-n tiles by O: one k-slice per tile instead
// of two with the hemisphere sign inside A, 186 DMMAs per quad instead of 316.
__host__ __device__ constexpr int MD2_NMAX(int m) { return 30 < 31 - m ? 30 : 31 - m; }
__host__ __device__ constexpr int MD2_NE(int m) { return (MD2_NMAX(m) / 2 + 1 + 7) / 8; }
__host__ __device__ constexpr int MD2_NO(int m) { return ((MD2_NMAX(m) + 1) / 2 + 7) / 8; }
__host__ __device__ constexpr int MD2_TOFF(int m) { int o = 0; for (int i = 0; i < m; i++) o += MD2_NE(i) + MD2_NO(i); return o; }
static_assert(MD2_TOFF(MX) == PD2_TTOT, "GlobTables::pq_dir2 layout");
template <int M> struct MdC2 {
    static constexpr int NT = MD2_NE(M) + MD2_NO(M);
    double cr[NT][2], ci[NT][2];
};
template <int M> __device__ __forceinline__ void md2_zero(MdC2<M> &c) {
#pragma unroll
    for (int i = 0; i < MdC2<M>::NT; i++) c.cr[i][0] = c.cr[i][1] = c.ci[i][0] = c.ci[i][1] = 0.0;
}
//   Aq : fragment table of this quad + lane ; Bl : slot + lane offset (row L%4, member L/4)
template <int M>
__device__ __forceinline__ void md2_mma(MdC2<M> &c, const double *__restrict__ Aq, const double *__restrict__ Bl) {
    constexpr int NE = MD2_NE(M), NT = MdC2<M>::NT;
    double a[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) a[i] = __ldg(Aq + (size_t)(MD2_TOFF(M) + i) * 32);
    const double br0 = Bl[(2 * M) * MQ_NM], br1 = Bl[4 * MD_RS + (2 * M) * MQ_NM];
    const double bi0 = Bl[(2 * M + 1) * MQ_NM], bi1 = Bl[4 * MD_RS + (2 * M + 1) * MQ_NM];
    const double er = br0 + br1, orr = br0 - br1, ei = bi0 + bi1, oi = bi0 - bi1;
#pragma unroll
    for (int i = 0; i < NT; i++) {
        dmma884(c.cr[i][0], c.cr[i][1], a[i], i < NE ? er : orr);
        dmma884(c.ci[i][0], c.ci[i][1], a[i], i < NE ? ei : oi);
    }
}
//   Xl : spectral output + lane offset (row 2 * (L/4) of a parity, members 2*(L%4), +1); all 32 rows n are written
//   sparse: the consumer reads the rows inside the nsh2 mask only (the model's spectral step: coefficients with
//   m + n <= 30 and their n +- 1 neighbours), so the exact zeros of the other rows (47 % of the field) are not written
template <int M>
__device__ __forceinline__ void md2_store(const MdC2<M> &c, double *__restrict__ Xl, const int col, const int sparse) {
    constexpr int NE = MD2_NE(M), NO = MD2_NO(M);
#pragma unroll
    for (int par = 0; par < 2; par++)
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (sparse && par + 2 * col + 16 * i > 31 - M) continue;  // vdspec reads n + 1 <= 31 - m (row (0,31) is a zero)
            double2 *pr = reinterpret_cast<double2 *>(Xl + (size_t)((2 * M) + M2 * (par + 16 * i)) * TILE);
            double2 *pi = reinterpret_cast<double2 *>(Xl + (size_t)((2 * M + 1) + M2 * (par + 16 * i)) * TILE);
            const bool have = i < (par ? NO : NE);
            const int q = have ? (par ? NE + i : i) : 0;
            // legendre.f90:206-218: rows outside the nsh2 mask are zero (the table holds 0 there as well)
            *pr = have ? make_double2(c.cr[q][0], c.cr[q][1]) : make_double2(0.0, 0.0);
            *pi = have ? make_double2(c.ci[q][0], c.ci[q][1]) : make_double2(0.0, 0.0);
        }
}

template <int LW>
__device__ __forceinline__ void g2s2_L(const Ctx &c, const FwdDesc *__restrict__ descs, const FwdOut *__restrict__ outs,
                                       const int nwork, const double *slots, const int lane, const int sparse) {
    constexpr int M3 = (LW != 7) ? 22 - LW : 15;
    const int kk = lane & 3, col = lane >> 2;
    const double *pq = c.G->pq_dir2 + lane;
    int g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        MdC2<LW> c0;
        MdC2<30 - LW> c1;
        MdC2<LW + 8> c2;
        MdC2<M3> c3;
        md2_zero(c0), md2_zero(c1), md2_zero(c2), md2_zero(c3);
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % G2_NS;
            const double *Bl = slots + sl * MD_SLOT + kk * MD_RS + col;
            const double *Aq = pq + (size_t)jq * (PD2_TTOT * 32);
            m2_sync(G2_FULL0 + sl, 512);
            md2_mma(c0, Aq, Bl), md2_mma(c1, Aq, Bl), md2_mma(c2, Aq, Bl);
            if (LW != 7) md2_mma(c3, Aq, Bl);
            m2_arrive(G2_EMPTY0 + sl, 512);
        }
        double *Xl = refp(c, t, outs[descs[f].fidx].dst, 0) + (size_t)(M2 * 2 * col) * TILE + MQ_NM * grp + 2 * kk;
        md2_store(c0, Xl, col, sparse), md2_store(c1, Xl, col, sparse), md2_store(c2, Xl, col, sparse);
        if (LW != 7) md2_store(c3, Xl, col, sparse);
    }
}

// stage-A loader: grid point i = 8q of item k sits in row q of the item's 12 rows (TMA box order)
template <int MODE> struct LdBox {
    const double *a, *b;
    double k0, sc;
    __device__ __forceinline__ double operator()(int i) const {
        const int r = (i >> 3) * 32;
        if (MODE == FM_PLAIN) return a[r];
        if (MODE == FM_COS) return a[r] * sc;
        if (MODE == FM_KE) {
            const double u = a[r], v = b[r];
            return 0.5 * (u * u + v * v);
        }
        if (MODE == FM_FLUXT) return (-a[r] * (b[r] - k0)) * sc;
        return (-a[r] * b[r]) * sc;  // FM_FLUX
    }
};
struct StSlotK {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int r, double v) const { p[r * MQ_NM] = v * sc; }
};

template <int MODE>
__device__ __forceinline__ void g2s2_F(const Ctx &c, const FwdDesc *__restrict__ descs, const int nwork, double *slots,
                                       double *bufs, const unsigned mbar0, const CUtensorMap *tmap, const int fw,
                                       const int lane) {
    constexpr int NOP = G2Cfg<MODE>::NOP, G2_NB = G2Cfg<MODE>::NB;
    const int jl = lane >> 3, mem = lane & 7;
    const bool issuer = (fw == 7 && lane == 0);  // warp 7 has no stage-B item
    const int npass = 2 * (IY / 4) * ((nwork - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
    // one thread requests the operand boxes of pass p (work item, quad, hemisphere decoded from p) into ring entry p % NB
    auto request = [&](int p) {
        const int item = p / (2 * (IY / 4)), rem = p - item * (2 * (IY / 4)), jq = rem >> 1, hemi = rem & 1;
        const int w = blockIdx.x + item * gridDim.x;
        const int grp = w & 3, t = (w >> 2) % c.ntiles, f = (w >> 2) / c.ntiles;
        const FwdDesc d = descs[f];
        const int lat0 = hemi ? 4 * jq : IL - 4 - 4 * jq;
        const unsigned mbar = mbar0 + 8 * (p % G2_NB);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(bufs + (size_t)(p % G2_NB) * NOP * M2_XH);
        const bool two = (MODE == FM_ALL) ? (d.mode >= FM_KE) : (NOP == 2);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((two ? 2 : 1) * M2_XH * 8) : "memory");
        const int ea = (int)((long long)t * c.scr_elems + (d.a & ~REF_SCR));
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
            "l"(tmap), "r"(mbar), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(ea)
            : "memory");
        if (two) {
            const int eb = (int)((long long)t * c.scr_elems + (d.b & ~REF_SCR));
            asm volatile(
                "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                    dst + M2_XH * 8),
                "l"(tmap), "r"(mbar), "r"(MQ_NM * grp), "r"(lat0), "r"(0), "r"(0), "r"(eb)
                : "memory");
        }
    };
    if (issuer)
        for (int p = 0; p < G2_NB - 1 && p < npass; p++) request(p);
    int p = 0, g = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int f = (w >> 2) / c.ntiles;
        const FwdDesc d = descs[f];
#pragma unroll 1
        for (int jq = 0; jq < IY / 4; jq++, g++) {
            const int sl = g % G2_NS;
#pragma unroll 1
            for (int hemi = 0; hemi < 2; hemi++, p++) {
                const int lat = (hemi ? 4 * jq : IL - 4 - 4 * jq) + jl, row = hemi ? 4 + jl : 3 - jl;
                double *ba = bufs + (size_t)(p % G2_NB) * NOP * M2_XH + lane;
                g2_mbar_wait(mbar0 + 8 * (p % G2_NB), (p / G2_NB) & 1);
                {   // stage A: item fw on rows 12fw .. 12fw+11, in place in operand buffer a
                    double *sa = ba + 12 * fw * 32;
                    const double sc = (d.kcos == 3) ? c_T.cosgr2[lat] : c_T.cosgr[lat];
                    if (MODE != FM_ALL) {
                        fftf_A0(LdBox<MODE>{sa, sa + (NOP == 2 ? M2_XH : 0), d.k0, sc}, sa);
                    } else {  // warp-uniform switch on the field's loader mode
                        if (d.mode == FM_PLAIN) fftf_A0(LdBox<FM_PLAIN>{sa, sa, d.k0, sc}, sa);
                        else if (d.mode == FM_COS) fftf_A0(LdBox<FM_COS>{sa, sa, d.k0, sc}, sa);
                        else if (d.mode == FM_KE) fftf_A0(LdBox<FM_KE>{sa, sa + M2_XH, d.k0, sc}, sa);
                        else if (d.mode == FM_FLUXT) fftf_A0(LdBox<FM_FLUXT>{sa, sa + M2_XH, d.k0, sc}, sa);
                        else fftf_A0(LdBox<FM_FLUX>{sa, sa + M2_XH, d.k0, sc}, sa);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                m2_sync(G2_GRP, 256);
                // every F warp has finished stage B of the previous pass: its ring entry is free for pass p + NB - 1
                if (issuer && p + G2_NB - 1 < npass) request(p + G2_NB - 1);
                if (hemi == 0 && g >= G2_NS) m2_sync(G2_EMPTY0 + sl, 512);
                double *S = slots + sl * MD_SLOT + row * MD_RS + mem;
                const StSlotK st{S, c_T.fc[3]};
                if (fw == 0) {
                    fftf_B0(ba, st), fftf_B6(ba, st);
                    S[MQ_NM] = 0.0;  // fourier.f90:117: Im of m = 0
                } else if (fw == 1) {
                    fftf_B1(ba, st);
                } else if (fw == 2) {
                    fftf_B2(ba, st);
                } else if (fw == 3) {
                    fftf_B3(ba, st);
                } else if (fw == 4) {
                    fftf_B4(ba, st);
                } else if (fw == 5) {
                    fftf_B5(ba, st);
                }
                // no proxy fence here: stage B only READS the ring entry (the next tensor load into it is requested after the
                // next 256-thread barrier) and writes the slot, which the async proxy never touches
                if (hemi == 1) m2_arrive(G2_FULL0 + sl, 512);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_grid2spec_mma2(const Ctx c, const FwdDesc *__restrict__ descs,
                                                           const FwdOut *__restrict__ outs, int nwork,
                                                           const __grid_constant__ CUtensorMap tmap, int sparse) {
    extern __shared__ __align__(128) double g2_sm[];
    constexpr int G2_NB = G2Cfg<MODE>::NB;
    double *bufs = g2_sm, *slots = g2_sm + (size_t)G2_NB * G2Cfg<MODE>::NOP * M2_XH;
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(slots + G2_NS * MD_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < G2_NB) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8 * threadIdx.x));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    switch (warp) {
        case 0: g2s2_L<0>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 1: g2s2_L<1>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 2: g2s2_L<2>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 3: g2s2_L<3>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 4: g2s2_L<4>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 5: g2s2_L<5>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 6: g2s2_L<6>(c, descs, outs, nwork, slots, lane, sparse); break;
        case 7: g2s2_L<7>(c, descs, outs, nwork, slots, lane, sparse); break;
        default: g2s2_F<MODE>(c, descs, nwork, slots, bufs, mbar0, &tmap, warp - 8, lane); break;
    }
}

template <int MODE>
static void launch_g2s2_mode(cudaStream_t s, const Ctx &c, const FwdDesc *d, const FwdOut *o, int nf, int sparse) {
    static int sms = 0;
    constexpr size_t SMEM = G2Cfg<MODE>::SMEM;
    static_assert(SMEM <= 232448, "shared memory per CTA on sm_100a");
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(k_grid2spec_mma2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) != cudaSuccess) {
            fprintf(stderr, "speedy_b200: k_grid2spec_mma2 needs %zu bytes of shared memory per CTA (sm_100a)\n", SMEM);
            abort();
        }
    }
    const int nwork = nf * c.ntiles * (TILE / MQ_NM);
    k_grid2spec_mma2<MODE><<<nwork < sms ? nwork : sms, 512, SMEM, s>>>(c, d, o, nwork, s2g2_tensor_map(c), sparse);
}
// all operand fields must live in the scratch arena (true for the model step's lists)
void launch_grid2spec_mma2(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, const FwdOut *o, int nf, int sparse) {
    if (!nf) return;
    switch (mode) {
        case FM_ALL: launch_g2s2_mode<FM_ALL>(s, c, d, o, nf, sparse); break;  // mixed list, FwdDesc::mode per field
        case FM_PLAIN: launch_g2s2_mode<FM_PLAIN>(s, c, d, o, nf, sparse); break;
        case FM_COS: launch_g2s2_mode<FM_COS>(s, c, d, o, nf, sparse); break;
        case FM_KE: launch_g2s2_mode<FM_KE>(s, c, d, o, nf, sparse); break;
        case FM_FLUXT: launch_g2s2_mode<FM_FLUXT>(s, c, d, o, nf, sparse); break;
        default: launch_g2s2_mode<FM_FLUX>(s, c, d, o, nf, sparse); break;
    }
}

}  // namespace spdy
