// Is the residual correction of the table-free RCPPS (devmath.cuh, RcpArith) ever needed on this GPU?
// For every bin k: d = 4097 + 2k, y = MUFU.RCP(d), q = rint(y * 2^25); compare with the table (q_ref = 4096 + tab16[k]).
// nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_rcp_check mufu_rcp_check.cu && ./mufu_rcp_check
#include <cstdio>
#include <cstdint>
static const uint16_t tab16_h[2048] = {
#include "../../lpcnet_b200/csrc/rcpps_table.inc"
};
__global__ void check(const uint16_t *tab, int *out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2048) return;
    const float d = (float)(4097 + 2 * k);
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    const float q = __fsub_rn(__fmaf_rn(y, 33554432.f, 12582912.f), 12582912.f);
    const int qi = (int)q, ref = 4096 + tab[k];
    out[k] = qi - ref;
    // distance of the MUFU quotient from the rounding boundary, in units of 1e-6
    out[2048 + k] = (int)(1e6f * fabsf(fabsf(__fmaf_rn(y, 33554432.f, -q)) - 0.5f));
}
__global__ void check_c(const uint16_t *tab, int *out, float C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2048) return;
    const float d = (float)(4097 + 2 * k);
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    const float q = __fsub_rn(__fmaf_rn(y, C, 12582912.f), 12582912.f);
    out[k] = (int)q - (4096 + tab[k]);
}
__global__ void check_m(const uint16_t *tab, int *out, float magic)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2048) return;
    const float d = (float)(4097 + 2 * k);
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    const int q = __float_as_int(__fmaf_rn(y, 33554432.f, magic)) - __float_as_int(magic);
    out[k] = q - (4096 + tab[k]);
}
int main()
{
    uint16_t *dt; int *dout; static int h[4096];
    cudaMalloc(&dt, sizeof(tab16_h)); cudaMalloc(&dout, sizeof(h));
    cudaMemcpy(dt, tab16_h, sizeof(tab16_h), cudaMemcpyHostToDevice);
    check<<<8, 256>>>(dt, dout);
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    // ties: an odd magic number rounds them to odd instead of even quotients
    for (int mg = 0; mg < 4; mg++) {
        static int h3[2048];
        check_m<<<8, 256>>>(dt, dout, 12582912.f + (float)mg);
        cudaMemcpy(h3, dout, sizeof(h3), cudaMemcpyDeviceToHost);
        int b = 0, first = -1; for (int k = 0; k < 2048; k++) if (h3[k]) { b++; if (first < 0) first = k; }
        printf("magic = 1.5*2^23 + %d: %d mismatches (first k=%d)\n", mg, b, first);
    }
    // candidate multipliers C = 2^25 + 4j (a relative bias of j * 2^-23): does one of them need no correction at all?
    for (int j = -4; j <= 12; j++) {
        static int h2[2048];
        check_c<<<8, 256>>>(dt, dout, 33554432.f + 4.f * j);
        cudaMemcpy(h2, dout, sizeof(h2), cudaMemcpyDeviceToHost);
        int b = 0, first = -1; for (int k = 0; k < 2048; k++) if (h2[k]) { b++; if (first < 0) first = k; }
        printf("C = 2^25 %+d: %d mismatches (first k=%d)\n", 4 * j, b, first);
    }
    int bad = 0, closest = 1 << 30;
    for (int k = 0; k < 2048; k++) { if (h[k]) { if (bad < 20) printf("k=%d diff=%d\n", k, h[k]); bad++; } if (h[2048 + k] < closest) closest = h[2048 + k]; }
    printf("mismatches without correction: %d of 2048; closest approach to a rounding boundary: %d e-6\n", bad, closest);
    return cudaGetLastError() != cudaSuccess;
}
