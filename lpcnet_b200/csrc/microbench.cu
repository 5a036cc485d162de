// microbench.cu — measured peak of the unit that bounds the per-sample kernel: the L1/shared-memory data pipe.
//
// The per-sample kernel keeps every weight in shared memory, so its roofline is shared-memory bandwidth, for which the
// driver-written MEASURED_PEAKS.json has no entry (it holds an HBM copy and a cuBLAS GEMM).  This file measures it the
// same way those peaks are measured — a trivially simple kernel that does nothing but the operation in question:
// every SM streams conflict-free LDS.128 (each warp-wide load = 512 contiguous bytes = 4 wavefronts) out of a 64 KB
// window with 8 independent loads in flight per thread and 32 resident warps.  Reported: GB/s over the whole chip
// (CUDA events) and bytes/clock/SM (clock64 inside the kernel), which is 128 on paper.
// bench.py calls it through lpcnet_b200_measure_smem_peak() and divides the kernel's algorithmic bytes/s by it.
#include <cstdint>
#include <cuda_runtime.h>
#include "engine.h"
#include "../../include/lpcnet_b200.h"

namespace lpcnet_b200 {

constexpr int MB_THREADS = 1024, MB_WINDOW = 64 * 1024, MB_UNROLL = 8;

template <int WIDTH>      // bytes per lane and load: 16 (LDS.128), 8 (LDS.64), 4 (LDS.32)
__global__ void __launch_bounds__(MB_THREADS, 1) smem_stream_kernel(int iters, uint32_t *sink, long long *cycles)
{
    extern __shared__ __align__(16) uint8_t win[];
    for (int i = threadIdx.x; i < MB_WINDOW / 4; i += MB_THREADS) reinterpret_cast<uint32_t *>(win)[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(win);
    uint32_t acc = 0;
    uint32_t off = threadIdx.x * WIDTH;                          // a warp reads 32 x WIDTH contiguous bytes: conflict-free
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        uint32_t v[MB_UNROLL][4] = {};
#pragma unroll
        for (int u = 0; u < MB_UNROLL; u++) {
            const uint32_t a = base + ((off + u * (MB_THREADS * WIDTH)) & (MB_WINDOW - 1));
            if (WIDTH == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "r"(a));
            else if (WIDTH == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[u][0]), "=r"(v[u][1]) : "r"(a));
            else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[u][0]) : "r"(a));
        }
#pragma unroll
        for (int u = 0; u < MB_UNROLL; u++) acc ^= v[u][0] ^ v[u][1] ^ v[u][2] ^ v[u][3];
        off += 32 * WIDTH * 5;                                   // move every warp's window (still contiguous per warp)
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;                       // keeps the loads alive
    if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
}

template <int WIDTH>
static int run_one(int sms, int iters, double *gbs, double *bytes_per_clk_sm)
{
    uint32_t *sink = nullptr; long long *cyc = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess || cudaMalloc(&cyc, sizeof(long long) * sms) != cudaSuccess) return -1;
    auto k = smem_stream_kernel<WIDTH>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_WINDOW);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, MB_THREADS, MB_WINDOW>>>(iters / 8 + 1, sink, cyc);       // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<<<sms, MB_THREADS, MB_WINDOW>>>(iters, sink, cyc);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -1;
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    long long *h = new long long[sms];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double cmax = 0; for (int i = 0; i < sms; i++) if ((double)h[i] > cmax) cmax = (double)h[i];
    delete[] h;
    const double bytes_cta = (double)iters * MB_UNROLL * MB_THREADS * WIDTH;
    *gbs = bytes_cta * sms / (best * 1e-3) / 1e9;
    *bytes_per_clk_sm = bytes_cta / cmax;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink); cudaFree(cyc);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace lpcnet_b200

using namespace lpcnet_b200;

// out[0..1] = LDS.128 {GB/s whole chip, bytes/clk/SM}, out[2..3] = LDS.64, out[4..5] = LDS.32, out[6] = SM count
extern "C" int lpcnet_b200_measure_smem_peak(int device, double *out)
{
    int sms = 0;
    if (!out || cudaSetDevice(device) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        set_error("measure_smem_peak: bad device %d", device); return -1;
    }
    const int iters = 3000;                                      // ~1.5 ms per launch at 128 B/clk/SM
    if (run_one<16>(sms, iters, &out[0], &out[1]) || run_one<8>(sms, iters * 2, &out[2], &out[3]) || run_one<4>(sms, iters * 4, &out[4], &out[5])) {
        set_error("measure_smem_peak: %s", cudaGetErrorString(cudaGetLastError())); return -1;
    }
    out[6] = sms;
    return 0;
}
