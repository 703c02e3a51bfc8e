// Micro-benchmark: cycles per whole-line 96-point inverse FFT (register variant, fft96_reg_gen.cuh) for one warp,
// inputs and outputs in shared memory, with 1 or 2 warps per scheduler.  Build (from pyspeedy_b200/csrc):
//   nvcc -O3 -std=c++17 -I. -gencode arch=compute_100a,code=sm_100a -o ../../tools/microbench/fftline ../../tools/microbench/fftline.cu
#include <cstdio>
#include "spdy.cuh"
namespace spdy {
#include "fft96_reg_gen.cuh"
struct LdS { const double *p; __device__ __forceinline__ double operator()(int r) const { return p[r * 32]; } };
struct StS { double *p; __device__ __forceinline__ void operator()(int i, double v) const { p[i * 32] = v; } };
__global__ void __launch_bounds__(256, 1) k(long long *cyc, double *out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *in = sm + warp * (62 + 96) * 32, *o = in + 62 * 32;
    for (int i = lane; i < 62 * 32; i += 32) in[i] = 1.0 / (1 + i);
    __syncthreads();
    const LdS ld{in + lane};
    const StS st{o + lane};
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double x[96];
        rfftb_A0(ld, x), rfftb_A1(ld, x), rfftb_A2(ld, x), rfftb_A3(ld, x), rfftb_A4(ld, x), rfftb_A5(ld, x), rfftb_A6(ld, x);
        rfftb_B0(x, st), rfftb_B1(x, st), rfftb_B2(x, st), rfftb_B3(x, st), rfftb_B4(x, st), rfftb_B5(x, st), rfftb_B6(x, st), rfftb_B7(x, st);
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    if (o[lane] == 1.2345) out[0] = o[lane];
}
}
int main() {
    long long *cyc; double *out;
    cudaMalloc(&cyc, 148 * 8 * 8); cudaMalloc(&out, 8);
    for (int warps = 4; warps <= 8; warps += 4) {
        const int smem = warps * (62 + 96) * 32 * 8;
        cudaFuncSetAttribute(spdy::k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        spdy::k<<<148, warps * 32, smem>>>(cyc, out, 200);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148 * 8];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%d warps/SM: %s  cycles per line-FFT per warp = %.0f\n", warps, cudaGetErrorString(e), (double)h[0] / 200);
    }
    return 0;
}
