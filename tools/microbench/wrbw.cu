// Micro-benchmark: sustained HBM write bandwidth of a pure store stream (8 GiB region, each byte written once)
// for warp stores of 256-byte rows and of 4 x 64-byte segments (the 8-member pattern of the fused transform),
// and for a mixed stream (one 256-byte read per two 256-byte writes).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o wrbw wrbw.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double *p, const double *q, size_t nrows) {  // rows of 32 doubles
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (size_t)gridDim.x * (blockDim.x >> 5);
    if (MODE == 0) {
        for (size_t r = warp; r < nrows; r += nw) p[r * 32 + lane] = (double)r;
    } else if (MODE == 1) {  // 4 rows x 8 members per instruction; the group index walks through the 4 segments
        for (size_t r4 = warp; r4 < nrows; r4 += nw) {
            const size_t blk = r4 / 4, grp = r4 % 4;  // rows 4*blk .. 4*blk+3, members 8*grp..
            p[(blk * 4 + (lane >> 3)) * 32 + grp * 8 + (lane & 7)] = (double)r4;
        }
    } else {
        double acc = 0;
        for (size_t r = warp; r < nrows; r += nw) {
            if ((r & 1) == 0) acc += q[(r >> 1) * 32 + lane];
            p[r * 32 + lane] = acc;
        }
    }
}
int main() {
    const size_t bytes = 8ull << 30, nrows = bytes / 256;
    double *p, *q;
    cudaMalloc(&p, bytes);
    cudaMalloc(&q, bytes / 2);
    cudaMemset(q, 0, bytes / 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const char *names[3] = {"256-byte rows", "4 x 64-byte segments", "256-byte rows + 1 read per 2 writes"};
    for (int m = 0; m < 3; m++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<148 * 8, 256>>>(p, q, nrows);
            else if (m == 1) k<1><<<148 * 8, 256>>>(p, q, nrows);
            else k<2><<<148 * 8, 256>>>(p, q, nrows);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %.3f ms  write %.0f GB/s%s\n", names[m], ms, bytes / (ms * 1e-3) / 1e9, m == 2 ? " (+ half as much read)" : "");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
