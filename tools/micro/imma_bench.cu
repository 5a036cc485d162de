// micro-benchmark: latency / throughput of legacy int8 mma.sync (IMMA.16832.U8.S8) on sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(long long *out, int *sink, int iters, int active_warps)
{
    const int warp = threadIdx.x >> 5;
    int c[CHAINS][4];
    for (int j = 0; j < CHAINS; j++) for (int i = 0; i < 4; i++) c[j][i] = threadIdx.x + i + j;
    uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 ^ 0x55, a3 = a0 + 7, b0 = 0x01ff02fe, b1 = 0x7f80017f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < active_warps) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < CHAINS; j++) imma(c[j], a0, a1, a2, a3, b0, b1);
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int j = 0; j < CHAINS; j++) for (int i = 0; i < 4; i++) s += c[j][i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main()
{
    long long *d; int *sink; cudaMalloc(&d, 8 * 256); cudaMalloc(&sink, 4 * 1024 * 256);
    const int iters = 4096;
    for (int aw : {1, 4, 8, 16}) {
        long long h;
        k<1><<<1, 512>>>(d, sink, iters, aw); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d chains=1: %.2f cycles per IMMA per warp (dependent)\n", aw, (double)h / iters);
        k<4><<<1, 512>>>(d, sink, iters, aw); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d chains=4: %.2f cycles per IMMA per warp; SM rate %.3f IMMA/cycle\n", aw, (double)h / iters / 4, aw * 4.0 * iters / (double)h);
    }
    // exactness: compare one IMMA against a scalar computation
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
