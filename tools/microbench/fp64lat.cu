// Micro-benchmark: dependent-issue latency of FP64 instructions on sm_100a (one warp, clock64 around a chain):
// DADD, DFMA, mma.sync.m8n8k4.f64 (DMMA), and DMMA issue interval with 2/4/8 independent chains.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64lat fp64lat.cu ; run: ./fp64lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int N = 2048;
template <int MODE, int CH> __global__ void k(const double *g, double *out, long long *cyc) {
    const int lane = threadIdx.x & 31;
    double a = g[lane], b = g[32 + lane];
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = g[64 + i];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N / 8; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) asm volatile("add.f64 %0, %0, %1;" : "+d"(c[i]) : "d"(a));
                else if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
                else dmma(c[2 * i], c[2 * i + 1], a, b);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int CH> void run(const char *name, const double *g, double *out, long long *cyc, int warps) {
    k<MODE, CH><<<1, 32 * warps>>>(g, out, cyc);
    cudaDeviceSynchronize();
    k<MODE, CH><<<1, 32 * warps>>>(g, out, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-6s chains=%d warps=%2d : %.1f cycles per instruction per warp (%.1f per chain step)\n", name, CH, warps,
           (double)h / (N * CH), (double)h / N);
}
int main() {
    double *g, *out;
    long long *cyc;
    cudaMalloc(&g, 4096 * 8), cudaMalloc(&out, 4096 * 8), cudaMalloc(&cyc, 64);
    cudaMemset(g, 0, 4096 * 8);
    run<0, 1>("DADD", g, out, cyc, 1), run<0, 4>("DADD", g, out, cyc, 1), run<0, 8>("DADD", g, out, cyc, 1);
    run<0, 8>("DADD", g, out, cyc, 4), run<0, 8>("DADD", g, out, cyc, 8), run<0, 8>("DADD", g, out, cyc, 16);
    run<1, 1>("DFMA", g, out, cyc, 1), run<1, 8>("DFMA", g, out, cyc, 1), run<1, 8>("DFMA", g, out, cyc, 8);
    run<2, 1>("DMMA", g, out, cyc, 1), run<2, 2>("DMMA", g, out, cyc, 1), run<2, 4>("DMMA", g, out, cyc, 1);
    run<2, 8>("DMMA", g, out, cyc, 1), run<2, 4>("DMMA", g, out, cyc, 4), run<2, 4>("DMMA", g, out, cyc, 8);
    run<2, 2>("DMMA", g, out, cyc, 8);
    return 0;
}
