// Micro-benchmark: FP64 tensor-core (mma.sync.m8n8k4.f64) versus FP64 FMA throughput on sm_100a, and the cost of
// feeding DMMA A-fragments from shared memory (one LDS.64 per lane per fragment).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu ; run: ./dmma
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ITER = 4096;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE> __global__ void __launch_bounds__(256) k(const double *__restrict__ g, double *out) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0.0;
    double a = g[lane], b = g[32 + lane];
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        if (MODE == 0) {  // 8 independent DMMA (8 x 256 FMA per warp)
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
        } else if (MODE == 1) {  // 16 independent DFMA chains x 4 (64 x 32 FMA per warp = 8 DMMA-equivalents)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fma(a, b, c[i]);
        } else {  // DMMA with a fresh A fragment from shared memory for every pair of DMMAs
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double aa = sm[((it * 8 + i) * 32 + lane) & 4095];
                dmma(c[2 * i], c[2 * i + 1], aa, b);
                dmma(c[2 * i + 2], c[2 * i + 3], aa, a);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 12345.678) out[threadIdx.x] = s;
}

template <int MODE> int run(const char *name, const double *g, double *out, int bps) {
    const int nb = 148 * bps;
    k<MODE><<<nb, 256>>>(g, out);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(g, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma_total = (double)nb * 8 * ITER * 8 * 256;  // warps x iters x 8 DMMA-equivalents x 256 FMA
    printf("%-34s blocks/SM=%d  %7.3f ms  %7.2f TFLOP/s\n", name, bps, ms, 2.0 * fma_total / (ms * 1e-3) / 1e12);
    return 0;
}

int main() {
    double *g, *out;
    CK(cudaMalloc(&g, 4096 * 8));
    CK(cudaMalloc(&out, 4096 * 8));
    CK(cudaMemset(g, 0, 4096 * 8));
    for (int b = 1; b <= 4; b *= 2) {
        run<0>("DMMA m8n8k4", g, out, b);
        run<1>("DFMA", g, out, b);
        run<2>("DMMA + LDS.64 A-fragment / 2 mma", g, out, b);
    }
    return 0;
}
