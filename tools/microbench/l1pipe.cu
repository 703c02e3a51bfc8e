// Micro-benchmark: cost (SM cycles per warp instruction) of the operand-delivery options for a warp-uniform table
// value on sm_100a: LDG/LDS broadcast of 64 and 128 bits, coalesced 64-bit rows, shuffle, constant-bank operands.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l1pipe l1pipe.cu ; run: ./l1pipe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ITER = 256, UNR = 16;
__constant__ double c_tab[4096];

template <int MODE> __global__ void __launch_bounds__(256) k(const double *__restrict__ g, double *out, long long *cyc) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        const int base = ((it * 7 + warp * 13) & 63) * 32;  // warp-uniform, varies
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            if (MODE == 0) {  // LDG.64 broadcast
                a0 += __ldg(g + base + u);
            } else if (MODE == 1) {  // LDG.128 broadcast
                const double2 v = __ldg(reinterpret_cast<const double2 *>(g + base) + u);
                a0 += v.x, a1 += v.y;
            } else if (MODE == 2) {  // LDS.64 broadcast
                a0 += sm[base + u];
            } else if (MODE == 3) {  // LDS.128 broadcast
                const double2 v = reinterpret_cast<const double2 *>(sm + base)[u];
                a0 += v.x, a1 += v.y;
            } else if (MODE == 4) {  // LDG.64 coalesced row (256 B)
                a0 += __ldg(g + base + u * 32 + lane);
            } else if (MODE == 5) {  // LDS.64 conflict-free row
                a0 += sm[((base + u * 32) & 4095) + lane];
            } else if (MODE == 6) {  // shuffle broadcast of a register value
                a0 += __shfl_sync(0xffffffffu, a1 + (double)u, u & 31);
            } else if (MODE == 7) {  // constant-bank operand, uniform runtime offset
                a0 = fma(c_tab[base + u], a1, a0);
            } else if (MODE == 8) {  // pure DFMA (reference)
                a0 = fma(a1, a2, a0), a1 = fma(a0, a3, a1);
            } else if (MODE == 9) {  // LDG.32 broadcast x2 (hi/lo words)
                const int lo = __ldg(reinterpret_cast<const int *>(g + base + u)), hi = __ldg(reinterpret_cast<const int *>(g + base + u) + 1);
                a0 += __hiloint2double(hi, lo);
            } else if (MODE == 10) {  // STG.64 coalesced row
                out[(size_t)(blockIdx.x * 8 + warp) * 4096 + ((base + u * 32) & 4095) + lane] = a0 + u;
            } else if (MODE == 11) {  // LDS.32 broadcast x2
                const int *p = reinterpret_cast<const int *>(sm + base + u);
                a0 += __hiloint2double(p[1], p[0]);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a0 + a1 + a2 + a3 == 12345.678) out[threadIdx.x] = a0;
}

template <int MODE> int run(const char *name, const double *g, double *out, long long *cyc, int blocks_per_sm) {
    const int nb = 148 * blocks_per_sm;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    k<MODE><<<nb, 256, 32768>>>(g, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256, 32768>>>(g, out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8];
    CK(cudaMemcpy(h, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < nb; i++) avg += h[i];
    avg /= nb;
    // SM cycles per warp-level instruction of the tested kind: blocks_per_sm*8 warps share the SM
    const double per = avg / ((double)ITER * UNR * 8 * blocks_per_sm);
    printf("%-28s blocks/SM=%d  %8.3f SM-cycles per warp-instr  (%.3f ms)\n", name, blocks_per_sm, per, ms);
    return 0;
}

int main() {
    double *g, *out;
    long long *cyc;
    CK(cudaMalloc(&g, 4096 * 8 * 2));
    CK(cudaMalloc(&out, (size_t)148 * 8 * 8 * 4096 * 8));
    CK(cudaMalloc(&cyc, 148 * 8 * 8));
    CK(cudaMemset(g, 0, 4096 * 8 * 2));
    for (int b = 2; b <= 4; b += 2) {
        run<8>("DFMA x2", g, out, cyc, b);
        run<0>("LDG.64 broadcast", g, out, cyc, b);
        run<1>("LDG.128 broadcast", g, out, cyc, b);
        run<9>("LDG.32 broadcast x2", g, out, cyc, b);
        run<2>("LDS.64 broadcast", g, out, cyc, b);
        run<3>("LDS.128 broadcast", g, out, cyc, b);
        run<11>("LDS.32 broadcast x2", g, out, cyc, b);
        run<4>("LDG.64 row (L1 hit)", g, out, cyc, b);
        run<5>("LDS.64 row", g, out, cyc, b);
        run<6>("SHFL.64 (2x SHFL.32)", g, out, cyc, b);
        run<7>("DFMA c[bank][uniform]", g, out, cyc, b);
        run<10>("STG.64 row", g, out, cyc, b);
    }
    return 0;
}
