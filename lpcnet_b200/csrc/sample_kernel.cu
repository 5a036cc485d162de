// sample_kernel.cu — the 16 kHz autoregressive loop as ONE persistent, warp-specialised sm_100a kernel.
//
// Replaces (reference file:line):
//   lpcnet_synthesize_tail_impl  src/lpcnet.c:235-271   (LPC prediction, u-law, de-emphasis, clamp/round)
//   run_sample_network           src/lpcnet.c:146-167
//   compute_gru_a_input          src/nnet.c:484-491     (cond + 3 embedding rows, left to right)
//   compute_sparse_gru           src/nnet.c:410-448  +  sparse_sgemv_accum8x4 (int8) src/vec_avx.h:790-858
//   compute_gruB                 src/nnet.c:326-372  +  sgemv_accum8x4 src/vec_avx.h:690-755
//   sample_mdense                src/nnet.c:163-214  +  kiss99_rand src/kiss99.c:59-81
//   lin2ulaw / ulaw2lin          src/common.h:37-58
//
// One CTA = 32 independent streams, one CTA per SM, NWC + NWP + 1 warps:
//
//   NWC compute warps own 48/NWC neuron groups (8 neurons x {z,r,h}) of GRU_A each.  The integer GEMVs S = W.q(h) of the
//                     32 streams are small GEMMs and run on the tensor cores: mma.sync m16n8k16 (u8 x s8 -> s32, exact, so
//                     the summation order is free): M = 16 streams, N = the 8 neurons of a row group, K = four 8x4 weight
//                     blocks ("quad") whose column blocks may be anywhere (block-sparse): the A fragment of lane (gid, t)
//                     is gathered from the quantised state with ONE LDS.128 (the words of streams gid, gid+8, gid+16,
//                     gid+24 for the column block of slot t), the B fragment is one LDS.32 of the quad's weights.  Two MMAs
//                     per quad cover the 32 streams.  The accumulator layout makes lane (gid, t) own neurons 2t, 2t+1 of
//                     the group for streams gid + 8j: the fp32 state of those 8 (stream, neuron) pairs lives in its
//                     registers for the whole launch and the activations are evaluated there.  The sums do not depend
//                     on the sampled excitation, so they are computed FIRST (overlapping the previous sample's sampler
//                     and this sample's gather) and the gathered input term is added when it arrives:
//                     acc = rne((bias + diag*h + gin)*16256) + S  is the same integer the reference gets.
//   NWP producer warps gather the GRU_A input term cond + E_sig[a] + E_pred[b] + E_exc[c] (compute_gru_a_input) for the 32
//                     streams, one gate at a time, with 512-byte contiguous LDG.128 (4 L1 lines per request instead of
//                     32 for a per-lane gather) into two [32][392] fp32 tiles that the compute lanes read conflict-free.
//    1 sampler warp   (lane == stream) runs the strictly serial tail (two KISS99 draws, 8-level sigmoid tree with sequential
//                     fp32 dot products, ulaw2lin, order-16 LPC filter, de-emphasis, lin2ulaw) for its 32 streams.
//
// Weights, su-biases, the upper dual_fc levels and the sampler tables are staged into shared memory once per launch
// by TMA bulk copies (cp.async.bulk + mbarrier).  Roles hand data over with named barriers (bar.arrive / bar.sync).
#include <cstdint>
#include "engine.h"
#include "devmath.cuh"

#ifndef LPCNET_GB
#define LPCNET_GB 3      // units (4 LDG.128 each) a producer lane keeps in flight
#endif

namespace lpcnet_b200 {

namespace {

// named barriers: who arrives (A) / who waits (S) and the thread count each one is armed with
enum {
    BAR_IDX = 1,     // A sampler, S producers : indices of the next sample are in idx_s                 (32 + 96)
    BAR_FULL0 = 2,   // A producers, S compute : tile 0 holds gate r (1st phase) / gate z (2nd phase)     (96 + 32 NWC)
    BAR_FULL1 = 3,   // A producers, S compute : tile 1 holds gate h                                     (96 + 384)
    BAR_EMPTY0 = 4,  // A compute, S producers : tile 0 (gate r) consumed, may be overwritten with gate z (384 + 96)
    BAR_X = 5,       // S compute              : new quantised GRU_A state complete                       (384)
    BAR_ACCB = 6,    // S compute              : GRU_B partial sums complete                              (384)
    BAR_HB = 7       // A compute, S sampler   : GRU_B state of this sample is in hBs                     (384 + 32)
};
constexpr int CNT_IDX = 32 + NWP * 32, CNT_FULL = (NWP + NWC) * 32, CNT_C = NWC * 32, CNT_HB = NWC * 32 + 32;

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMA bulk copy global -> shared, completion on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ int4 lds128(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(addr));
    return v;
}
// D[16 streams][8 neurons] += A[16 streams][16 inputs] (u8) . B[16 inputs][8 neurons] (s8): exact int32
__device__ __forceinline__ void imma16816(int &c0, int &c1, int &c2, int &c3, uint32_t a0, uint32_t a1, uint32_t b0)
{
    asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3) : "r"(a0), "r"(a1), "r"(b0));
}
// acc[2j+i] += sum over `nq` quads, for stream gid+8j and neuron 2t+i of the row group.
//   w    : shared address of the first quad's weights + lane*4      (B fragment word of this lane)
//   meta : shared address of the first quad's meta + t*2            (xs_offset of slot t's column block)
//   xs   : shared address of the state buffer; the lane's 16-byte vector of slot t is at xs + (meta entry ^ gid*16)
// Software-pipelined: the meta entry of quad q+2 and the operands of quad q+1 are in flight while quad q is multiplied
// (the image keeps two quads of readable slack behind every list).
__device__ __forceinline__ void mma_quads(int (&acc)[8], uint32_t w, uint32_t meta, int nq, uint32_t xs, uint32_t gid16)
{
    uint32_t e1 = lds16(meta + QUAD_META_BYTES);
    int4 x = lds128(xs + (lds16(meta) ^ gid16));
    uint32_t wv = lds32(w);
    for (int q = 0; q < nq; q++) {
        const int4 xn = lds128(xs + (e1 ^ gid16));
        const uint32_t wn = lds32(w + QUAD_BYTES);
        e1 = lds16(meta + 2 * QUAD_META_BYTES);
        imma16816(acc[0], acc[1], acc[2], acc[3], (uint32_t)x.x, (uint32_t)x.y, wv);
        imma16816(acc[4], acc[5], acc[6], acc[7], (uint32_t)x.z, (uint32_t)x.w, wv);
        x = xn; wv = wn;
        w += QUAD_BYTES; meta += QUAD_META_BYTES;
    }
}

// Producer warps: ONE gate's input term for all 32 streams of the CTA,
//   G[s][k] = ((cond[s][k] + E_sig[a_s][k]) + E_pred[b_s][k]) + E_exc[c_s][k]        (nnet.c:484-491, left to right)
// Producer p serves streams p, p+NWP, ...: per stream the four row pointers are formed once and the 384 columns of the
// gate are covered by three 512-byte LDG.128 per row (4 L1 lines per request), i.e. 12 independent loads in flight per
// lane, then 12 fp32 adds and three 512-byte conflict-free STS.128 into the [32][388] tile.
__device__ __forceinline__ void gather_slice(float *__restrict__ G, const float *__restrict__ cond_f, int n, int cta_s0,
                                             const float *__restrict__ emb_sig, const float *__restrict__ emb_pred,
                                             const float *__restrict__ emb_exc, const int *__restrict__ idx_s,
                                             int gate, int p, int lane)
{
    const int col = gate * NA + lane * 4;
    for (int ss = p; ss < STREAMS_PER_CTA; ss += NWP) {
        const int sg = min(cta_s0 + ss, n - 1);
        const float *c = cond_f + (size_t)sg * (3 * NA) + col;
        const float *e0 = emb_sig + idx_s[ss] * (3 * NA) + col;
        const float *e1 = emb_pred + idx_s[32 + ss] * (3 * NA) + col;
        const float *e2 = emb_exc + idx_s[64 + ss] * (3 * NA) + col;
        float4 a[3], b[3], d[3], e[3];
#pragma unroll
        for (int j = 0; j < 3; j++) { a[j] = ldg4(c + 128 * j); b[j] = ldg4(e0 + 128 * j); d[j] = ldg4(e1 + 128 * j); e[j] = ldg4(e2 + 128 * j); }
        float *g = G + ss * GIN_ROW + lane * 4;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float4 r;
            r.x = __fadd_rn(__fadd_rn(__fadd_rn(a[j].x, b[j].x), d[j].x), e[j].x);
            r.y = __fadd_rn(__fadd_rn(__fadd_rn(a[j].y, b[j].y), d[j].y), e[j].y);
            r.z = __fadd_rn(__fadd_rn(__fadd_rn(a[j].z, b[j].z), d[j].z), e[j].z);
            r.w = __fadd_rn(__fadd_rn(__fadd_rn(a[j].w, b[j].w), d[j].w), e[j].w);
            *reinterpret_cast<float4 *>(g + 128 * j) = r;
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(SAMPLE_THREADS, 1) lpcnet_sample_kernel(const __grid_constant__ SampleParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemLayout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = P.n_streams;
    const int cta_s0 = blockIdx.x * STREAMS_PER_CTA;
    const int s_raw = cta_s0 + lane;
    const bool live = s_raw < n;
    const int s = live ? s_raw : n - 1;          // dead lanes shadow the last stream (all loads valid), stores masked

    // ---- stage the constant image with TMA bulk copies ----
    const uint32_t bar = smem_u32(smem + SM_MBAR);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, L.image_bytes);
        const uint32_t CH = 16384;
        for (uint32_t o = 0; o < L.image_bytes; o += CH) {
            const uint32_t nbytes = min(CH, L.image_bytes - o);
            bulk_g2s(smem_u32(smem + SM_IMAGE + o), P.image + o, nbytes, bar);
        }
    }

    const uint32_t *rcp = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_RCP);
    uint8_t *xs = smem + SM_XS;
    uint32_t *xbw = reinterpret_cast<uint32_t *>(smem + SM_XB);   // quantised GRU_B state, [4 words][32 lanes], lane == stream
    int *accB = reinterpret_cast<int *>(smem + SM_ACCB);
    float *hBs = reinterpret_cast<float *>(smem + SM_HBS);
    int *idx_s = reinterpret_cast<int *>(smem + SM_IDX);
    float *tile0 = reinterpret_cast<float *>(smem + SM_T0);
    float *tile1 = reinterpret_cast<float *>(smem + SM_T1);
    const int spf = P.spf;

    if (warp < NWC) {
        // =====================================================  compute warps  =====================================================
        mbar_wait(bar, 0);
        const int gid = lane >> 2, t = lane & 3;                         // MMA fragment coordinates of this lane
        const uint32_t gid16 = gid * 16;
        const uint32_t *grpA = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_GRPA);
        const uint32_t *dirA = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_DIRA) + warp * GPW * 3 * 2;
        const float *parA = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_PARA) + warp * GPW * 3 * 16 + 2 * t;
        const uint32_t metaA = smem_u32(smem + L.metaA) + t * 2;
        const uint32_t wA = smem_u32(smem + L.wA) + lane * 4;
        const uint32_t *dirB = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_DIRB);
        const uint32_t metaB = smem_u32(smem + L.metaB) + t * 2;
        const uint32_t wB = smem_u32(smem + L.wB) + lane * 4;
        const float *parB = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_PARB);
        const uint8_t *wBrec = smem + SM_IMAGE + IM_WBREC;
        const uint32_t xs0 = smem_u32(xs);
        // the 4 streams of this lane (rows gid, gid+8 of the two MMAs); dead streams shadow the last one, stores masked
        int sj[4]; bool livej[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int r = cta_s0 + gid + 8 * j; livej[j] = r < n; sj[j] = livej[j] ? r : n - 1; }

        int gcol[GPW];                                                   // tile column / neuron index of this lane's first neuron: 8*g + 2t
        uint32_t xoff[GPW];                                              // byte offset (in a state buffer) of this lane's two quantised neurons, stream gid
        float h[GPW][8];                                                 // fp32 state: [stream j][neuron i] at 2j+i
#pragma unroll
        for (int sl = 0; sl < GPW; sl++) {
            const int g = (int)grpA[warp * GPW + sl];
            gcol[sl] = 8 * g + 2 * t;
            xoff[sl] = xs_offset(2 * g + (t >> 1), gid) + (t & 1) * 2;
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int i = 0; i < 2; i++) h[sl][2 * j + i] = P.hA[(size_t)(gcol[sl] + i) * n + sj[j]];
        }
        // GRU_B neurons finished by this warp (lane == stream there): jb = warp + k*NWC < NB
        float hb[NBW];
#pragma unroll
        for (int k = 0; k < NBW; k++) hb[k] = P.hB[(size_t)min(warp + k * NWC, NB - 1) * n + s];
        // quantised copies of the restored state: xs <- q(hA), xb[0] <- q(hB)
#pragma unroll
        for (int sl = 0; sl < GPW; sl++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                *reinterpret_cast<uint16_t *>(xs + xoff[sl] + 4 * j) = (uint16_t)(quant_u8(h[sl][2 * j]) | (quant_u8(h[sl][2 * j + 1]) << 8));
#pragma unroll
        for (int k = 0; k < NBW; k++) {
            const int jb = warp + k * NWC;
            if (jb < NB) reinterpret_cast<uint8_t *>(xbw)[((jb >> 2) * 32 + lane) * 4 + (jb & 3)] = (uint8_t)quant_u8(hb[k]);
        }
        bar_sync(BAR_X, CNT_C);                                          // restored quantised state visible to all compute warps

        int step = 0;
        for (int f = 0; f < P.nframes; f++) {
            const float *condBp = P.condB + ((size_t)f * n + s) * (3 * NB);
            float cbz[NBW], cbr[NBW], cbh[NBW];
#pragma unroll
            for (int k = 0; k < NBW; k++) {
                const int jb = min(warp + k * NWC, NB - 1);
                cbz[k] = __ldg(condBp + jb); cbr[k] = __ldg(condBp + NB + jb); cbh[k] = __ldg(condBp + 2 * NB + jb);
            }
            for (int t_ = 0; t_ < spf; t_++, step++) {
                const int cur = step & 1, nxt = cur ^ 1;                 // double buffers of the quantised states
                const uint32_t xs_cur = xs0 + cur * XS_BYTES;
                uint8_t *xs_nxt = xs + nxt * XS_BYTES;
                int Sh[GPW][8];                                          // candidate-gate GEMV sums; later (bit pattern) rec_h * r, then h~
                // ---- A: candidate-gate GEMV  S_h = W_h . q(h)   (needs only the previous state: overlaps sampler + gather) ----
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const uint32_t *dir = dirA + sl * 3 * 2;
#pragma unroll
                    for (int i = 0; i < 8; i++) Sh[sl][i] = 0;
                    mma_quads(Sh[sl], wA + dir[4] * QUAD_BYTES, metaA + dir[4] * QUAD_META_BYTES, (int)dir[5], xs_cur, gid16);
                }
                // ---- B: reset gate r (nnet.c:431-435): GEMV first, then the gathered input term from tile 0 ----
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const float *par = parA + sl * 3 * 16;
                    const uint32_t *dir = dirA + sl * 3 * 2;
                    int Sr[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) Sr[i] = 0;
                    mma_quads(Sr, wA + dir[2] * QUAD_BYTES, metaA + dir[2] * QUAD_META_BYTES, (int)dir[3], xs_cur, gid16);
                    if (sl == 0) bar_sync(BAR_FULL0, CNT_FULL);          // gate r of all 32 streams is in tile 0
                    const float2 br = *reinterpret_cast<const float2 *>(par + 16), dr = *reinterpret_cast<const float2 *>(par + 24);
                    const float2 bh = *reinterpret_cast<const float2 *>(par + 32), dh = *reinterpret_cast<const float2 *>(par + 40);
                    const float bri[2] = {br.x, br.y}, dri[2] = {dr.x, dr.y}, bhi[2] = {bh.x, bh.y}, dhi[2] = {dh.x, dh.y};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 gv = *reinterpret_cast<const float2 *>(tile0 + (gid + 8 * j) * GIN_ROW + gcol[sl]);
                        const float gin[2] = {gv.x, gv.y};
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const float hv = h[sl][2 * j + i];
                            const int acc = acc_init(__fadd_rn(__fadd_rn(bri[i], __fmul_rn(dri[i], hv)), gin[i])) + Sr[2 * j + i];
                            const float r = sigmoid_approx(acc_finish(acc), rcp);
                            // candidate pre-activation: rec_h = bias + diag*h (+ GEMV), no input term (nnet.c:436-440); keep rec_h * r
                            const int acch = acc_init(__fadd_rn(bhi[i], __fmul_rn(dhi[i], hv))) + Sh[sl][2 * j + i];
                            Sh[sl][2 * j + i] = __float_as_int(__fmul_rn(acc_finish(acch), r));
                        }
                    }
                }
                bar_arrive(BAR_EMPTY0, CNT_FULL);                        // tile 0 may now receive gate z
                // ---- C: h~ = tanh(rec_h*r + gin_h) from tile 1 (nnet.c:443-445) ----
                bar_sync(BAR_FULL1, CNT_FULL);                           // gate h is in tile 1
#pragma unroll
                for (int sl = 0; sl < GPW; sl++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 gv = *reinterpret_cast<const float2 *>(tile1 + (gid + 8 * j) * GIN_ROW + gcol[sl]);
                        Sh[sl][2 * j] = __float_as_int(tanh_approx(__fadd_rn(__int_as_float(Sh[sl][2 * j]), gv.x), rcp));
                        Sh[sl][2 * j + 1] = __float_as_int(tanh_approx(__fadd_rn(__int_as_float(Sh[sl][2 * j + 1]), gv.y), rcp));
                    }
                // ---- D: update gate z (GEMV + input term from tile 0, 2nd phase), h <- z*h + (1-z)*h~ (nnet.c:446-447), new state ----
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const float *par = parA + sl * 3 * 16;
                    const uint32_t *dir = dirA + sl * 3 * 2;
                    int Sz[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) Sz[i] = 0;
                    mma_quads(Sz, wA + dir[0] * QUAD_BYTES, metaA + dir[0] * QUAD_META_BYTES, (int)dir[1], xs_cur, gid16);
                    if (sl == 0) bar_sync(BAR_FULL0, CNT_FULL);          // gate z of all 32 streams is in tile 0
                    const float2 bz = *reinterpret_cast<const float2 *>(par), dz = *reinterpret_cast<const float2 *>(par + 8);
                    const float bzi[2] = {bz.x, bz.y}, dzi[2] = {dz.x, dz.y};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 gv = *reinterpret_cast<const float2 *>(tile0 + (gid + 8 * j) * GIN_ROW + gcol[sl]);
                        const float gin[2] = {gv.x, gv.y};
                        uint32_t q[2];
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const float hv = h[sl][2 * j + i];
                            const int acc = acc_init(__fadd_rn(__fadd_rn(bzi[i], __fmul_rn(dzi[i], hv)), gin[i])) + Sz[2 * j + i];
                            const float z = sigmoid_approx(acc_finish(acc), rcp);
                            const float hn = __fadd_rn(__fmul_rn(z, hv), __fmul_rn(__fsub_rn(1.f, z), __int_as_float(Sh[sl][2 * j + i])));
                            h[sl][2 * j + i] = hn;
                            q[i] = quant_u8(hn);
                        }
                        // other buffer: readers of the old state are unaffected
                        *reinterpret_cast<uint16_t *>(xs_nxt + xoff[sl] + 4 * j) = (uint16_t)(q[0] | (q[1] << 8));
                    }
                }
                bar_sync(BAR_X, CNT_C);                                  // new quantised GRU_A state complete (tile 1 is dead: accB/hBs may use it)

                // ---------------- E: GRU_B input GEMV (48 x 384 int8, dense): warp = (row group, K part) ----------------
                if (warp < NWB) {
                    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const uint32_t q0 = dirB[warp * 2], nq = dirB[warp * 2 + 1];
                    mma_quads(acc, wB + q0 * QUAD_BYTES, metaB + q0 * QUAD_META_BYTES, (int)nq, xs0 + nxt * XS_BYTES, gid16);
                    const int rgp = warp / KPARTS, part = warp % KPARTS;
                    int *dst = accB + (part * 3 * NB + rgp * 8 + 2 * t) * ACCB_ROW + gid;
#pragma unroll
                    for (int j = 0; j < 4; j++) { dst[8 * j] = acc[2 * j]; dst[ACCB_ROW + 8 * j] = acc[2 * j + 1]; }
                }
                // ---------------- GRU_B finish (nnet.c:346-371), lane == stream: neurons warp, warp + NWC, ... ----------------
                {
                    const uint32_t *xbc = xbw + cur * 4 * 32;
                    const uint32_t xw[4] = {xbc[lane], xbc[32 + lane], xbc[64 + lane], xbc[96 + lane]};
                    uint8_t *xbn = reinterpret_cast<uint8_t *>(xbw + nxt * 4 * 32);
                    // recurrent side first: it only needs the previous GRU_B state, so it runs while other warps finish their partial sums
                    int rz[NBW], rr[NBW], rh[NBW];
#pragma unroll
                    for (int k2 = 0; k2 < NBW; k2++) {
                        const int jb = min(warp + k2 * NWC, NB - 1);
                        rz[k2] = acc_init(parB[3 * NB + jb]); rr[k2] = acc_init(parB[4 * NB + jb]); rh[k2] = acc_init(parB[5 * NB + jb]);
#pragma unroll
                        for (int k = 0; k < 4; k++) {    // W_rec layout [out/8][in/4][8][4] (dump_lpcnet.py:58-59)
                            rz[k2] = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + (((jb >> 3) * 4 + k) * 8 + (jb & 7)) * 4), rz[k2]);
                            rr[k2] = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + ((((NB + jb) >> 3) * 4 + k) * 8 + ((NB + jb) & 7)) * 4), rr[k2]);
                            rh[k2] = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + ((((2 * NB + jb) >> 3) * 4 + k) * 8 + ((2 * NB + jb) & 7)) * 4), rh[k2]);
                        }
                    }
                    bar_sync(BAR_ACCB, CNT_C);                           // all K-part partial sums are in accB
#pragma unroll
                    for (int k2 = 0; k2 < NBW; k2++) {
                        const int jb = warp + k2 * NWC;
                        if (jb >= NB) break;
                        int az = acc_init(__fadd_rn(parB[jb], cbz[k2])), ar = acc_init(__fadd_rn(parB[NB + jb], cbr[k2])), ah = acc_init(__fadd_rn(parB[2 * NB + jb], cbh[k2]));
#pragma unroll
                        for (int kp = 0; kp < KPARTS; kp++) {
                            az += accB[(kp * 3 * NB + jb) * ACCB_ROW + lane];
                            ar += accB[(kp * 3 * NB + NB + jb) * ACCB_ROW + lane];
                            ah += accB[(kp * 3 * NB + 2 * NB + jb) * ACCB_ROW + lane];
                        }
                        const float zz = sigmoid_approx(__fadd_rn(acc_finish(az), acc_finish(rz[k2])), rcp);
                        const float rrr = sigmoid_approx(__fadd_rn(acc_finish(ar), acc_finish(rr[k2])), rcp);
                        const float hh = tanh_approx(__fadd_rn(acc_finish(ah), __fmul_rn(acc_finish(rh[k2]), rrr)), rcp);
                        hb[k2] = __fadd_rn(__fmul_rn(zz, hb[k2]), __fmul_rn(__fsub_rn(1.f, zz), hh));
                        hBs[jb * 32 + lane] = hb[k2];
                        xbn[((jb >> 2) * 32 + lane) * 4 + (jb & 3)] = (uint8_t)quant_u8(hb[k2]);
                    }
                }
                __threadfence_block();
                bar_arrive(BAR_HB, CNT_HB);                              // GRU_B state of this sample is in hBs
            }
        }
        // ---- save the recurrent state ----
#pragma unroll
        for (int sl = 0; sl < GPW; sl++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (livej[j]) {
                    P.hA[(size_t)gcol[sl] * n + sj[j]] = h[sl][2 * j];
                    P.hA[(size_t)(gcol[sl] + 1) * n + sj[j]] = h[sl][2 * j + 1];
                }
        if (live) {
#pragma unroll
            for (int k = 0; k < NBW; k++) if (warp + k * NWC < NB) P.hB[(size_t)(warp + k * NWC) * n + s] = hb[k];
        }
    } else if (warp < NWC + NWP) {
        // =====================================================  producer warps  =====================================================
        const int p = warp - NWC;
        for (int f = 0; f < P.nframes; f++) {
            const float *condA_f = P.condA + (size_t)f * n * (3 * NA);
            for (int t = 0; t < spf; t++) {
                bar_sync(BAR_IDX, CNT_IDX);                              // indices of this sample are in idx_s
                gather_slice(tile0, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 1, p, lane);   // gate r
                __threadfence_block();
                bar_arrive(BAR_FULL0, CNT_FULL);
                gather_slice(tile1, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 2, p, lane);   // gate h
                __threadfence_block();
                bar_arrive(BAR_FULL1, CNT_FULL);
                bar_sync(BAR_EMPTY0, CNT_FULL);                          // gate r consumed
                gather_slice(tile0, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 0, p, lane);   // gate z
                __threadfence_block();
                bar_arrive(BAR_FULL0, CNT_FULL);
            }
        }
    } else {
        // =====================================================  sampler warp  =====================================================
        mbar_wait(bar, 0);
        const float *logit = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_LOGIT);
        const float *u2l = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_U2L);
        const float *fcw = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_FCW);

        float ls[LPC_ORDER], lpc[LPC_ORDER];
#pragma unroll
        for (int j = 0; j < LPC_ORDER; j++) ls[j] = P.last_sig[(size_t)j * n + s];
        float deemph = P.deemph[s];
        int last_exc = P.last_exc[s];
        Kiss99 rng;
        rng.z = P.rng[s]; rng.w = P.rng[(size_t)n + s]; rng.jsr = P.rng[2 * (size_t)n + s]; rng.jcong = P.rng[3 * (size_t)n + s];
        short *pcm_out = P.pcm + (size_t)s * P.pcm_stream_stride;

        for (int f = 0; f < P.nframes; f++) {
            {   // frame f uses the LPC computed from the features of frame f-2 (lpcnet.c:110-112), weighted by gamma^i (freq.c:299-308)
                const float *lp = P.lpc_raw + ((size_t)f * n + s) * LPC_ORDER;
                const float4 a = ldg4(lp), b = ldg4(lp + 4), c = ldg4(lp + 8), d = ldg4(lp + 12);
                const float raw[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
                for (int j = 0; j < LPC_ORDER; j++) lpc[j] = __fmul_rn(raw[j], __ldg(&P.gamma_pow[j]));
            }
            for (int t = 0; t < spf; t++) {
                // prediction and conditioning indices of this sample (lpcnet.c:251-254)
                float pred = 0.f;
#pragma unroll
                for (int j = 0; j < LPC_ORDER; j++) pred = __fsub_rn(pred, __fmul_rn(ls[j], lpc[j]));
                idx_s[lane] = lin2ulaw(ls[0]);
                idx_s[32 + lane] = lin2ulaw(pred);
                idx_s[64 + lane] = last_exc;
                __threadfence_block();
                bar_arrive(BAR_IDX, CNT_IDX);
                // thresholds (nnet.c:178-184): two RNG words -> 8 logits; does not depend on the network
                float thr[8];
                {
                    uint32_t r0 = kiss99_rand(rng), r1 = kiss99_rand(rng);
                    thr[0] = logit[r0 & 0xFF]; thr[1] = logit[(r0 >> 8) & 0xFF]; thr[2] = logit[(r0 >> 16) & 0xFF]; thr[3] = logit[r0 >> 24];
                    thr[4] = logit[r1 & 0xFF]; thr[5] = logit[(r1 >> 8) & 0xFF]; thr[6] = logit[(r1 >> 16) & 0xFF]; thr[7] = logit[r1 >> 24];
                }
                bar_sync(BAR_HB, CNT_HB);                                // wait for GRU_B
                float hbv[NB];
#pragma unroll
                for (int j = 0; j < NB; j++) hbv[j] = hBs[j * 32 + lane];
                int val = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) {                            // sample_mdense, nnet.c:186-211
                    const int i = (1 << b) | val;
                    float sum1, sum2, fac1, fac2;
                    // two sequential 16-term chains (one per channel), fed 8 weights at a time to keep the live set small
                    if (b < 6) {                                         // nodes < 64: rows in shared memory
                        const float *wr = fcw + i * FCW_ROW;
                        const float4 bf = *reinterpret_cast<const float4 *>(wr + 32);
                        sum1 = bf.x; sum2 = bf.y; fac1 = bf.z; fac2 = bf.w;
#pragma unroll
                        for (int j0 = 0; j0 < NB; j0 += 8) {
                            const float4 a0 = *reinterpret_cast<const float4 *>(wr + j0), a1 = *reinterpret_cast<const float4 *>(wr + j0 + 4);
                            const float4 c0 = *reinterpret_cast<const float4 *>(wr + NB + j0), c1 = *reinterpret_cast<const float4 *>(wr + NB + j0 + 4);
                            const float wa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, wc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                sum1 = __fadd_rn(sum1, __fmul_rn(wa[j], hbv[j0 + j]));
                                sum2 = __fadd_rn(sum2, __fmul_rn(wc[j], hbv[j0 + j]));
                            }
                        }
                    } else {                                             // lower levels: one 144-byte row per lane from global (L2-resident)
                        const float *wr = P.fcw + i * FCW_ROW;
                        const float4 bf = ldg4(wr + 32);
                        sum1 = bf.x; sum2 = bf.y; fac1 = bf.z; fac2 = bf.w;
#pragma unroll
                        for (int j0 = 0; j0 < NB; j0 += 8) {
                            const float4 a0 = ldg4(wr + j0), a1 = ldg4(wr + j0 + 4), c0 = ldg4(wr + NB + j0), c1 = ldg4(wr + NB + j0 + 4);
                            const float wa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, wc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                sum1 = __fadd_rn(sum1, __fmul_rn(wa[j], hbv[j0 + j]));
                                sum2 = __fadd_rn(sum2, __fmul_rn(wc[j], hbv[j0 + j]));
                            }
                        }
                    }
                    sum1 = __fmul_rn(fac1, tanh_approx(sum1, rcp));
                    sum2 = __fmul_rn(fac2, tanh_approx(sum2, rcp));
                    sum1 = __fadd_rn(sum1, sum2);
                    val = (val << 1) | (thr[b] < sum1 ? 1 : 0);
                }
                const int exc = val;
                float pcm = __fadd_rn(pred, u2l[exc]);                   // lpcnet.c:260
#pragma unroll
                for (int j = LPC_ORDER - 1; j > 0; j--) ls[j] = ls[j - 1];
                ls[0] = pcm;
                last_exc = exc;
                pcm = __fadd_rn(pcm, __fmul_rn(0.85f, deemph));          // PREEMPH, lpcnet.c:265
                deemph = pcm;
                if (pcm < -32767) pcm = -32767;
                if (pcm > 32767) pcm = 32767;
                if (live) pcm_out[(size_t)f * spf + t] = (short)__double2int_rd(0.5 + (double)pcm);   // (int)floor(.5 + pcm)
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < LPC_ORDER; j++) P.last_sig[(size_t)j * n + s] = ls[j];
            P.deemph[s] = deemph;
            P.last_exc[s] = last_exc;
            P.rng[s] = rng.z; P.rng[(size_t)n + s] = rng.w; P.rng[2 * (size_t)n + s] = rng.jsr; P.rng[3 * (size_t)n + s] = rng.jcong;
        }
    }
}

int sample_kernel_smem_ok(uint32_t bytes)
{
    return bytes <= 227u * 1024u;
}

cudaError_t launch_sample_kernel(const SampleParams &p, cudaStream_t st)
{
    // per-device attribute; cheap enough to set on every launch (one launch covers >= 160 x n_streams samples)
    cudaError_t e = cudaFuncSetAttribute(lpcnet_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const int ctas = (p.n_streams + STREAMS_PER_CTA - 1) / STREAMS_PER_CTA;
    lpcnet_sample_kernel<<<ctas, SAMPLE_THREADS, p.L.total_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace lpcnet_b200
