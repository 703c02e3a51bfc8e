// speedy-b200: definitions shared by the two fused transform kernels (fused_mma3.cu: spectral -> grid,
// fused_mma2.cu: grid -> spectral).  Work item of both = (field, tile, group of MQ_NM = 8 members); the Fourier rows of a
// latitude block travel between the Legendre (DMMA) warps and the FFT warps in shared-memory slots; grid rows enter and
// leave the SM as boxes of one 5-D TMA tensor map of the scratch arena.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace spdy {

constexpr int MQ_NM = 8;                      // members per work item
constexpr int MQ_RS = M2 * MQ_NM + 8;         // spec -> grid slot row stride (doubles): 62 x 64 B + 64 B -> rows 2i, 2i+1 in
                                              // different bank halves: C stores and F row loads conflict-free
constexpr int MD_RS = M2 * MQ_NM + 4;         // grid -> spec slot row stride: 62 x 64 B + 32 B -> the four rows of a B fragment
                                              // fall into different bank quarters (conflict-free LDS.64)
constexpr int MD_SLOT = 8 * MD_RS;            // doubles per grid -> spec slot (latitude quad, both hemispheres)
constexpr int M2_XH = IX * 32;                // doubles per FFT exchange buffer / TMA box (32 lines = 4 latitudes x 8 members)

__device__ __forceinline__ void m2_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void m2_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// FP64 tensor-core MMA (the only FP64 tensor path on sm_100a: tcgen05 has no f64 kind): C[8x8] += A[8x4] * B[4x8]
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
struct LdSlot {  // FFT stage-A loader: Fourier row r of a slot, this thread's member
    const double *p;
    __device__ __forceinline__ double operator()(int r) const { return p[r * MQ_NM]; }
};
// stage-B outputs go back IN PLACE into the exchange rows the thread has just read (item k: rows 12k..12k+11 hold
// grid points k + 8q, q = 0..11, afterwards), scaled for kcos = 2 (fourier.f90:88-92)
struct StExchK {
    double *p;
    double sc;
    __device__ __forceinline__ void operator()(int i, double v) const { p[(i >> 3) * 32] = v * sc; }
};

// The scratch arena as a 5-D FP64 tensor for the TMA loads and stores: element (lane, lat, q, k, e) lives at
//   scr + 8 * lane + 256 * (e + 96 * lat + 8 * q + k)   bytes,
// i.e. grid point i = k + 8q of latitude row lat of the (96,48) field that starts at tile-relative element offset e
// (lane stride 8 B, point stride 256 B: the 32-member tile layout of spdy.cuh).  Box = [8][4][12][8][1].
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// nlat = latitude rows per box: 4 (quad per hemisphere); 8 (octet per hemisphere) is used by tools/experiments/fused_mma5.cu only
static const CUtensorMap &s2g2_tensor_map(const Ctx &c, const int nlat = 4) {
    static EncodeTiledFn encode = nullptr;
    static CUtensorMap maps[2];
    static const void *ks_scr[2] = {nullptr, nullptr};
    static long long ks_elems[2] = {-1, -1};
    static int ks_tiles[2] = {-1, -1};
    const int slot = nlat == 8 ? 1 : 0;
    CUtensorMap &map = maps[slot];
    const void *&k_scr = ks_scr[slot];
    long long &k_elems = ks_elems[slot];
    int &k_tiles = ks_tiles[slot];
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
        const cuuint32_t box[5] = {MQ_NM, (cuuint32_t)nlat, 12, 8, 1}, estr[5] = {1, 1, 1, 1, 1};
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

}  // namespace spdy
