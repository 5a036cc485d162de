// sample_kernel.cu — the 16 kHz autoregressive loop as ONE persistent sm_100a kernel.
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
// Mapping (DESIGN.md "sample kernel"): one CTA = 32 independent streams, LANE == STREAM.  All 32 lanes of a warp
// execute the same (warp-uniform) walk over the block-sparse weights, so every weight word is fetched from shared
// memory ONCE per 32 streams (broadcast LDS.128) instead of once per stream, and no cross-lane reduction exists:
// each lane finishes the 8 outputs of a block row group in its own registers with dp4a.u32.s32 (one dp4a == one
// block row, integer-exact like maddubs+madd under the WeightClip pair constraint).  16 compute warps each own 3
// neuron groups (8 neurons x {z,r,h}) of GRU_A — fp32 state lives in registers for the whole utterance, only the
// quantised u8 state is exchanged through shared memory — plus one neuron of GRU_B.  A 17th warp runs the strictly
// serial tail (tree sampler, LPC filter, u-law, de-emphasis) for its 32 streams.  Weights, su-biases, dual_fc and the
// sampler tables are staged into shared memory once per launch by TMA bulk copies (cp.async.bulk + mbarrier).
#include <cstdint>
#include "engine.h"
#include "devmath.cuh"

namespace lpcnet_b200 {

namespace {

enum { BAR_IDX = 1, BAR_X = 2, BAR_ACCB = 3, BAR_HB = 4, BAR_GF = 5, BAR_GE = 6 };   // GF: gather tile full, GE: tile free again

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

// Cooperative gather of ONE gate's GRU_A input for the CTA's 32 streams (compute_gru_a_input, nnet.c:484-491):
//   G[s][k] = ((cond[s][k] + E_sig[a_s][k]) + E_pred[b_s][k]) + E_exc[c_s][k],   k in [0, 384)
// Warp w serves streams 2w and 2w+1; its 32 lanes read consecutive float4 of the four rows (512 B contiguous per
// instruction = 4 L1 lines, instead of 32 lines for a per-lane gather) and store the sums transposed-by-row into
// the shared tile, from which every lane (== stream) later reads its own 8-neuron slices conflict-free.
__device__ __forceinline__ void gather_gate(float *__restrict__ G, const float *__restrict__ cond_f, int n, int cta_s0,
                                            const float *__restrict__ emb_sig, const float *__restrict__ emb_pred,
                                            const float *__restrict__ emb_exc, const int *__restrict__ idx_s,
                                            int gate, int warp, int lane)
{
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int ss = 2 * warp + k;
        const int sg = min(cta_s0 + ss, n - 1);
        const float *c = cond_f + (size_t)sg * (3 * NA) + gate * NA + lane * 4;
        const float *e0 = emb_sig + (size_t)idx_s[ss] * (3 * NA) + gate * NA + lane * 4;
        const float *e1 = emb_pred + (size_t)idx_s[32 + ss] * (3 * NA) + gate * NA + lane * 4;
        const float *e2 = emb_exc + (size_t)idx_s[64 + ss] * (3 * NA) + gate * NA + lane * 4;
        float *g = G + ss * GIN_ROW + lane * 4;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float4 a = ldg4(c + 128 * j), b = ldg4(e0 + 128 * j), d = ldg4(e1 + 128 * j), e = ldg4(e2 + 128 * j);
            float4 r;
            r.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, b.x), d.x), e.x);
            r.y = __fadd_rn(__fadd_rn(__fadd_rn(a.y, b.y), d.y), e.y);
            r.z = __fadd_rn(__fadd_rn(__fadd_rn(a.z, b.z), d.z), e.z);
            r.w = __fadd_rn(__fadd_rn(__fadd_rn(a.w, b.w), d.w), e.w);
            *reinterpret_cast<float4 *>(g + 128 * j) = r;
        }
    }
}
__device__ __forceinline__ void load_gin(float gin[8], const float *__restrict__ G, int lane, int g)
{
    const float4 a = *reinterpret_cast<const float4 *>(G + lane * GIN_ROW + 8 * g);
    const float4 b = *reinterpret_cast<const float4 *>(G + lane * GIN_ROW + 8 * g + 4);
    gin[0] = a.x; gin[1] = a.y; gin[2] = a.z; gin[3] = a.w; gin[4] = b.x; gin[5] = b.y; gin[6] = b.z; gin[7] = b.w;
}

// acc[r] += sum over `nb` (even) 8x4 blocks; weights broadcast from smem, activations one word per lane
__device__ __forceinline__ void gemv_blocks(int acc[8], const uint8_t *__restrict__ w, const uint16_t *__restrict__ meta,
                                            int nb, const uint8_t *__restrict__ xs_lane)
{
    for (int b = 0; b < nb; b += 2) {
        const uint32_t m = *reinterpret_cast<const uint32_t *>(meta + b);
        const uint32_t x0 = *reinterpret_cast<const uint32_t *>(xs_lane + (m & 0xFFFFu));
        const uint32_t x1 = *reinterpret_cast<const uint32_t *>(xs_lane + (m >> 16));
        const int4 wa = *reinterpret_cast<const int4 *>(w);
        const int4 wb = *reinterpret_cast<const int4 *>(w + 16);
        const int4 wc = *reinterpret_cast<const int4 *>(w + 32);
        const int4 wd = *reinterpret_cast<const int4 *>(w + 48);
        acc[0] = dp4a_us(x0, wa.x, acc[0]); acc[1] = dp4a_us(x0, wa.y, acc[1]);
        acc[2] = dp4a_us(x0, wa.z, acc[2]); acc[3] = dp4a_us(x0, wa.w, acc[3]);
        acc[4] = dp4a_us(x0, wb.x, acc[4]); acc[5] = dp4a_us(x0, wb.y, acc[5]);
        acc[6] = dp4a_us(x0, wb.z, acc[6]); acc[7] = dp4a_us(x0, wb.w, acc[7]);
        acc[0] = dp4a_us(x1, wc.x, acc[0]); acc[1] = dp4a_us(x1, wc.y, acc[1]);
        acc[2] = dp4a_us(x1, wc.z, acc[2]); acc[3] = dp4a_us(x1, wc.w, acc[3]);
        acc[4] = dp4a_us(x1, wd.x, acc[4]); acc[5] = dp4a_us(x1, wd.y, acc[5]);
        acc[6] = dp4a_us(x1, wd.z, acc[6]); acc[7] = dp4a_us(x1, wd.w, acc[7]);
        w += 64;
    }
}

}  // namespace

__global__ void __launch_bounds__(SAMPLE_THREADS, 1) lpcnet_sample_kernel(const __grid_constant__ SampleParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemLayout &L = P.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = P.n_streams;
    const int s_raw = blockIdx.x * STREAMS_PER_CTA + lane;
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
    uint32_t *xbw = reinterpret_cast<uint32_t *>(smem + SM_XB);
    int *accB = reinterpret_cast<int *>(smem + SM_ACCB);
    float *hBs = reinterpret_cast<float *>(smem + SM_HBS);
    int *idx_s = reinterpret_cast<int *>(smem + SM_IDX);
    float *gin_tile = reinterpret_cast<float *>(smem + SM_GIN);
    const int cta_s0 = blockIdx.x * STREAMS_PER_CTA;
    const int spf = P.spf;

    if (warp < NWC) {
        // =====================================================  compute warps  =====================================================
        mbar_wait(bar, 0);
        const uint32_t *grpA = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_GRPA);
        const uint32_t *dirA = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_DIRA) + warp * GPW * 3 * 2;
        const float *parA = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_PARA) + warp * GPW * 3 * 16;
        const uint16_t *metaA = reinterpret_cast<const uint16_t *>(smem + L.metaA);
        const uint8_t *wA = smem + L.wA;
        const uint32_t *dirB = reinterpret_cast<const uint32_t *>(smem + SM_IMAGE + IM_DIRB);
        const uint16_t *metaB = reinterpret_cast<const uint16_t *>(smem + L.metaB);
        const uint8_t *wB = smem + L.wB;
        const float *parB = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_PARB);
        const uint8_t *wBrec = smem + SM_IMAGE + IM_WBREC;

        int grp[GPW];
        float h[GPW][8];
#pragma unroll
        for (int sl = 0; sl < GPW; sl++) {
            grp[sl] = (int)grpA[warp * GPW + sl];
#pragma unroll
            for (int i = 0; i < 8; i++) h[sl][i] = P.hA[(size_t)(8 * grp[sl] + i) * n + s];
        }
        const int jb = warp;                               // the GRU_B neuron this warp finishes (NB == NWC)
        float hb = P.hB[(size_t)jb * n + s];
        // quantised copies of the restored state: xs[0] <- q(hA), xb[0] <- q(hB)
#pragma unroll
        for (int sl = 0; sl < GPW; sl++) {
            uint32_t w0 = quant_u8(h[sl][0]) | (quant_u8(h[sl][1]) << 8) | (quant_u8(h[sl][2]) << 16) | (quant_u8(h[sl][3]) << 24);
            uint32_t w1 = quant_u8(h[sl][4]) | (quant_u8(h[sl][5]) << 8) | (quant_u8(h[sl][6]) << 16) | (quant_u8(h[sl][7]) << 24);
            reinterpret_cast<uint32_t *>(xs)[(2 * grp[sl]) * 32 + lane] = w0;
            reinterpret_cast<uint32_t *>(xs)[(2 * grp[sl] + 1) * 32 + lane] = w1;
        }
        reinterpret_cast<uint8_t *>(xbw)[((jb >> 2) * 32 + lane) * 4 + (jb & 3)] = (uint8_t)quant_u8(hb);

        int step = 0;
        for (int f = 0; f < P.nframes; f++) {
            const float *condA_f = P.condA + (size_t)f * n * (3 * NA);
            const float *condBp = P.condB + ((size_t)f * n + s) * (3 * NB);
            const float cbz = __ldg(condBp + jb), cbr = __ldg(condBp + NB + jb), cbh = __ldg(condBp + 2 * NB + jb);
            for (int t = 0; t < spf; t++, step++) {
                const int cur = step & 1, nxt = cur ^ 1;
                bar_sync(BAR_IDX, SAMPLE_THREADS);                       // indices of this step are in idx_s
                const uint8_t *xs_cur = xs + cur * XS_BYTES + lane * 4;
                uint32_t *xs_nxt = reinterpret_cast<uint32_t *>(xs + nxt * XS_BYTES);
                float rg[GPW][8];                                        // reset gate, then (in place) the candidate h~

                // ---------------- GRU_A, gate by gate: cooperative gather -> tile -> owners consume ----------------
                // reset gate r: rec = bias + diag*h + gin (nnet.c:431-435), + int8 GEMV, sigmoid
                gather_gate(gin_tile, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 1, warp, lane);
                bar_sync(BAR_GF, NWC * 32);
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const float *par = parA + sl * 3 * 16;
                    const uint32_t *dir = dirA + sl * 3 * 2;
                    float gin[8]; int acc[8];
                    load_gin(gin, gin_tile, lane, grp[sl]);
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] = acc_init(__fadd_rn(__fadd_rn(par[16 + i], __fmul_rn(par[24 + i], h[sl][i])), gin[i]));
                    gemv_blocks(acc, wA + (size_t)dir[2] * 32, metaA + dir[2], (int)dir[3], xs_cur);
#pragma unroll
                    for (int i = 0; i < 8; i++) rg[sl][i] = sigmoid_approx(acc_finish(acc[i]), rcp);
                }
                bar_sync(BAR_GE, NWC * 32);
                // candidate: rec = bias + diag*h (no input term, nnet.c:436-440); h~ = tanh(rec*r + gin_h) (:443-445)
                gather_gate(gin_tile, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 2, warp, lane);
                bar_sync(BAR_GF, NWC * 32);
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const float *par = parA + sl * 3 * 16;
                    const uint32_t *dir = dirA + sl * 3 * 2;
                    float gin[8]; int acc[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] = acc_init(__fadd_rn(par[32 + i], __fmul_rn(par[40 + i], h[sl][i])));
                    gemv_blocks(acc, wA + (size_t)dir[4] * 32, metaA + dir[4], (int)dir[5], xs_cur);
                    load_gin(gin, gin_tile, lane, grp[sl]);
#pragma unroll
                    for (int i = 0; i < 8; i++) rg[sl][i] = tanh_approx(__fadd_rn(__fmul_rn(acc_finish(acc[i]), rg[sl][i]), gin[i]), rcp);
                }
                bar_sync(BAR_GE, NWC * 32);
                // update gate z, then h <- z*h + (1-z)*h~ (nnet.c:446-447) and the new quantised state
                gather_gate(gin_tile, condA_f, n, cta_s0, P.emb_sig, P.emb_pred, P.emb_exc, idx_s, 0, warp, lane);
                bar_sync(BAR_GF, NWC * 32);
#pragma unroll
                for (int sl = 0; sl < GPW; sl++) {
                    const int g = grp[sl];
                    const float *par = parA + sl * 3 * 16;
                    const uint32_t *dir = dirA + sl * 3 * 2;
                    float gin[8]; int acc[8];
                    load_gin(gin, gin_tile, lane, g);
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] = acc_init(__fadd_rn(__fadd_rn(par[i], __fmul_rn(par[8 + i], h[sl][i])), gin[i]));
                    gemv_blocks(acc, wA + (size_t)dir[0] * 32, metaA + dir[0], (int)dir[1], xs_cur);
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float z = sigmoid_approx(acc_finish(acc[i]), rcp);
                        const float hn = __fadd_rn(__fmul_rn(z, h[sl][i]), __fmul_rn(__fsub_rn(1.f, z), rg[sl][i]));
                        h[sl][i] = hn;
                        q[i] = quant_u8(hn);
                    }
                    xs_nxt[(2 * g) * 32 + lane] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
                    xs_nxt[(2 * g + 1) * 32 + lane] = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
                }
                bar_sync(BAR_X, NWC * 32);                               // new quantised GRU_A state complete

                // ---------------- GRU_B input GEMV (48 x 384 int8): warp = (row group, K half) ----------------
                if (warp < NWB) {
                    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    const uint32_t b0 = dirB[warp * 2], nb = dirB[warp * 2 + 1];
                    gemv_blocks(acc, wB + (size_t)b0 * 32, metaB + b0, (int)nb, xs + nxt * XS_BYTES + lane * 4);
                    const int rg = warp >> 1, half = warp & 1;
#pragma unroll
                    for (int i = 0; i < 8; i++) accB[(half * 3 * NB + rg * 8 + i) * 32 + lane] = acc[i];
                }
                bar_sync(BAR_ACCB, NWC * 32);

                // ---------------- GRU_B finish: this warp's neuron jb (nnet.c:346-371) ----------------
                {
                    const uint32_t *xbc = xbw + cur * 4 * 32;
                    const uint32_t x0 = xbc[lane], x1 = xbc[32 + lane], x2 = xbc[64 + lane], x3 = xbc[96 + lane];
                    // su-biases and recurrent weight words of rows jb, 16+jb, 32+jb; W_rec layout [out/8][in/4][8][4] (dump_lpcnet.py:58-59)
                    int az = acc_init(__fadd_rn(parB[jb], cbz)) + accB[jb * 32 + lane] + accB[(3 * NB + jb) * 32 + lane];
                    int ar = acc_init(__fadd_rn(parB[NB + jb], cbr)) + accB[(NB + jb) * 32 + lane] + accB[(3 * NB + NB + jb) * 32 + lane];
                    int ah = acc_init(__fadd_rn(parB[2 * NB + jb], cbh)) + accB[(2 * NB + jb) * 32 + lane] + accB[(3 * NB + 2 * NB + jb) * 32 + lane];
                    int rz = acc_init(parB[3 * NB + jb]), rr = acc_init(parB[4 * NB + jb]), rh = acc_init(parB[5 * NB + jb]);
                    const uint32_t xw[4] = {x0, x1, x2, x3};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        rz = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + (((jb >> 3) * 4 + k) * 8 + (jb & 7)) * 4), rz);
                        rr = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + ((((NB + jb) >> 3) * 4 + k) * 8 + ((NB + jb) & 7)) * 4), rr);
                        rh = dp4a_us(xw[k], *reinterpret_cast<const int *>(wBrec + ((((2 * NB + jb) >> 3) * 4 + k) * 8 + ((2 * NB + jb) & 7)) * 4), rh);
                    }
                    const float zz = sigmoid_approx(__fadd_rn(acc_finish(az), acc_finish(rz)), rcp);
                    const float rrr = sigmoid_approx(__fadd_rn(acc_finish(ar), acc_finish(rr)), rcp);
                    const float hh = tanh_approx(__fadd_rn(acc_finish(ah), __fmul_rn(acc_finish(rh), rrr)), rcp);
                    hb = __fadd_rn(__fmul_rn(zz, hb), __fmul_rn(__fsub_rn(1.f, zz), hh));
                    hBs[jb * 32 + lane] = hb;
                    reinterpret_cast<uint8_t *>(xbw + nxt * 4 * 32)[((jb >> 2) * 32 + lane) * 4 + (jb & 3)] = (uint8_t)quant_u8(hb);
                }
                __threadfence_block();
                bar_arrive(BAR_HB, SAMPLE_THREADS);                      // GRU_B state of this step is in hBs
            }
        }
        // ---- save the recurrent state ----
        if (live) {
#pragma unroll
            for (int sl = 0; sl < GPW; sl++)
#pragma unroll
                for (int i = 0; i < 8; i++) P.hA[(size_t)(8 * grp[sl] + i) * n + s] = h[sl][i];
            P.hB[(size_t)jb * n + s] = hb;
        }
    } else {
        // =====================================================  sampler warp  =====================================================
        mbar_wait(bar, 0);
        const float *logit = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_LOGIT);
        const float *u2l = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_U2L);
        const float *fcw = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_FCW);
        const float *fcb = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_FCB);
        const float *fcf = reinterpret_cast<const float *>(smem + SM_IMAGE + IM_FCF);
        short *pcm_s = reinterpret_cast<short *>(smem + SM_PCM);

        float ls[LPC_ORDER], lpc[LPC_ORDER];
#pragma unroll
        for (int j = 0; j < LPC_ORDER; j++) ls[j] = P.last_sig[(size_t)j * n + s];
        float deemph = P.deemph[s];
        int last_exc = P.last_exc[s];
        Kiss99 rng;
        rng.z = P.rng[s]; rng.w = P.rng[(size_t)n + s]; rng.jsr = P.rng[2 * (size_t)n + s]; rng.jcong = P.rng[3 * (size_t)n + s];

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
                bar_arrive(BAR_IDX, SAMPLE_THREADS);
                // thresholds (nnet.c:178-184): two RNG words -> 8 logits; does not depend on the network
                float thr[8];
                {
                    uint32_t r0 = kiss99_rand(rng), r1 = kiss99_rand(rng);
                    thr[0] = logit[r0 & 0xFF]; thr[1] = logit[(r0 >> 8) & 0xFF]; thr[2] = logit[(r0 >> 16) & 0xFF]; thr[3] = logit[r0 >> 24];
                    thr[4] = logit[r1 & 0xFF]; thr[5] = logit[(r1 >> 8) & 0xFF]; thr[6] = logit[(r1 >> 16) & 0xFF]; thr[7] = logit[r1 >> 24];
                }
                bar_sync(BAR_HB, SAMPLE_THREADS);                        // wait for GRU_B
                float hbv[NB];
#pragma unroll
                for (int j = 0; j < NB; j++) hbv[j] = hBs[j * 32 + lane];
                int val = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) {                            // sample_mdense, nnet.c:186-211
                    const int i = (1 << b) | val;
                    const float *wr = fcw + i * FCW_ROW;
                    float sum1 = fcb[i], sum2 = fcb[256 + i];
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        sum1 = __fadd_rn(sum1, __fmul_rn(wr[j], hbv[j]));
                        sum2 = __fadd_rn(sum2, __fmul_rn(wr[NB + j], hbv[j]));
                    }
                    sum1 = __fmul_rn(fcf[i], tanh_approx(sum1, rcp));
                    sum2 = __fmul_rn(fcf[256 + i], tanh_approx(sum2, rcp));
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
                pcm_s[lane * PCM_ROW + t] = (short)__double2int_rd(0.5 + (double)pcm);   // (int)floor(.5 + pcm)
            }
            // ---- flush the frame's PCM tile: 32 streams x spf samples, coalesced along time ----
            __syncwarp();
            for (int ss = 0; ss < STREAMS_PER_CTA; ss++) {
                const int sg = blockIdx.x * STREAMS_PER_CTA + ss;
                if (sg >= n) break;
                short *dst = P.pcm + (size_t)sg * P.pcm_stream_stride + (size_t)f * spf;
                for (int t = lane; t < spf; t += 32) dst[t] = pcm_s[ss * PCM_ROW + t];
            }
            __syncwarp();
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
