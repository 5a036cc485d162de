// engine.h — internal structures shared by the host-side model loader and the sm_100a kernels.
// Not part of the C ABI (include/*.h is).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace lpcnet_b200 {

// ---- model dimensions (default LPCNet: training_tf2/lpcnet.py:234; generated nnet_data.h) ----
constexpr int NA = 384;            // GRU_A units
constexpr int NB = 16;             // GRU_B units
constexpr int COND = 128;          // conditioning width
constexpr int PITCH_EMBED = 64;
constexpr int NB_FEAT = 20;
constexpr int FRAME_IN = NB_FEAT + PITCH_EMBED;   // 84
constexpr int LPC_ORDER = 16;
constexpr int NB_BANDS = 18;
constexpr int FRAME_SIZE = 160;
constexpr int FEATURES_DELAY = 2;
constexpr int WINDOW_SIZE = 320;
constexpr int FREQ_SIZE = 161;

// ---- per-sample kernel geometry (warp-specialised CTA) ----
constexpr int STREAMS_PER_CTA = 32;          // stream slots of a CTA; the launcher may leave some dead (SampleParams::spc) to spread a small batch over all SMs
#ifndef LPCNET_NWC
#define LPCNET_NWC 16
#endif
constexpr int NWC = LPCNET_NWC;              // compute warps (12, 16 or 24): each owns NGRP/NWC neuron groups of GRU_A, a (row group, K part) of the
                                             // GRU_B input GEMV and up to ceil(16/NWC) GRU_B neurons.  24 warps x 2 groups keeps the per-thread
                                             // working set small (h, S_h, S_z = 48 registers) so that 7 warps per scheduler hide the LDS latency
#ifndef LPCNET_NWP
#define LPCNET_NWP 6
#endif
constexpr int NWP = LPCNET_NWP;              // producer warps: cooperative gather of the GRU_A input rows.  The gather is latency-bound (L2 hits, ~1k cycles
                                             // under load), so what matters is loads in flight: NWP warps x (registers/4) LDG.128 each
constexpr int NGRP = NA / 8;                 // 48 groups of 8 neurons (one 8-row block group per gate)
constexpr int GPW = NGRP / NWC;              // neuron groups per compute warp
static_assert(NGRP % NWC == 0, "compute warps must divide the 48 neuron groups");
constexpr int SAMPLE_THREADS = (NWC + NWP + 2) * 32;   // + 2 sampler warps, one per half (tree sampler, LPC filter, u-law, de-emphasis)
constexpr int XS_BYTES = (NA / 4) * 32 * 4;  // quantised GRU_A state of 32 streams: [96 column blocks][32 words], see xs_offset()
constexpr int FCW_ROW = 36;                  // dual_fc row: 32 weights (16 per channel) + {bias0, bias1, factor0, factor1}; 144 B stride = 16 mod 128, so the
                                             // per-lane LDS.128 row reads of lanes on different nodes mostly land in different 4-bank groups
constexpr int FCW_SMEM_NODES = 64;           // at most the tree levels 0..5 (nodes 1..63) live in shared memory, the rest is read from global (L2);
                                             // a model whose block lists need the room keeps fewer (32, 16, 8: model.cu), the count travels in the image
constexpr int KPARTS = (NWC >= 24) ? 4 : 2;  // K split of the GRU_B input GEMV
constexpr int NWB = 6 * KPARTS;              // warps used by the GRU_B input GEMV: (row group 0..5) x (K part)
static_assert(NWB <= NWC, "one (row group, K part) of GRU_B per compute warp");
constexpr int HALF = 16;                     // the 32 streams of a CTA are stepped as two halves of 16 (= the M of one MMA), half a sample apart:
                                             // while the sampler / gather of one half run, the compute warps work on the other half
constexpr int NFIN = NB * HALF / 32;         // compute warps that finish GRU_B for a half: lane = (neuron parity, stream in half)
static_assert(NFIN <= NWC, "GRU_B finishing warps");
constexpr int ACCB_ROW = 20;                 // int32 per output row of the GRU_B partial sums (16 streams + 4 pad: the MMA accumulator stores of a warp
                                             // (rows 2t, columns gid) then fall into 32 different banks)

// ---- int8 flavour: the integer GEMVs run on the tensor cores (mma.sync m16n8k16, u8 x s8 -> s32, exact) ----
// A "quad" is four 8x4 weight blocks of one 8-row group = one MMA: 16 streams x (4 blocks x 4 inputs) x 8 outputs.
//   weights: 128 B per quad, word [gid][t] = the 4 int8 weights of output row gid for the block in slot t   (B fragment of lane gid*4+t)
//   meta   : 4 x u16 per quad, one per slot: xs_offset(column block of the slot, stream 0)
// Quantised state xs: column block c (inputs 4c..4c+3) is a 128-byte row = 2 halves x 8 streams x 2 words; the word of
// stream s = 16*H + 8*jj + gid sits at
//   c*128 + (((H << 6) | (gid << 3)) ^ (((c & 1) << 6) | ((c & 2) << 4))) + jj*4
// i.e. the two streams {gid, gid+8} of a half that lane (gid, t) feeds to its MMA are one 8-byte vector, and the XOR
// spreads the four slots of a quad (whose column blocks are ordered to have distinct c & 3 where possible) over the
// four 32-byte bank groups that a half-warp (gid 0..3 or 4..7) does not already separate: conflict-free LDS.64.
constexpr uint32_t QUAD_BYTES = 128, QUAD_META_BYTES = 8;
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t xs_offset(uint32_t c, uint32_t s)
{
    return c * 128u + (((((s >> 4) & 1u) << 6) | ((s & 7u) << 3)) ^ (((c & 1u) << 6) | ((c & 2u) << 4))) + ((s >> 3) & 1u) * 4u;
}

// ---- shared-memory map of the per-sample kernel ----
// Everything whose size does not depend on the model's sparsity pattern sits at a COMPILE-TIME offset (keeps the
// addresses out of registers); only the four block-sparse arrays are placed at run-time offsets behind them.
// [SM_IMAGE, SM_IMAGE + image_bytes) is copied verbatim from the global "SMEM image" built at model-load time
// (TMA bulk copies); [0, SM_IMAGE) is the mutable working set.
constexpr uint32_t al128(uint32_t x) { return (x + 127u) & ~127u; }
constexpr int GIN_ROW = 392;                                       // floats per stream in a gather tile: 384 + 8 pad => row stride = 8 words mod 32:
                                                                   // the LDS.64 of lanes (gid, t) = row gid, column 2t hit 32 different banks per half-warp
constexpr int NTILE = 4;                                           // ring of gather tiles, filled in the order (half A: r, z, h), (half B: r, z, h), ...
constexpr uint32_t TILE_BYTES = HALF * GIN_ROW * 4;                // one gate of one half: float [16 streams][GIN_ROW]
constexpr uint32_t SM_XS    = 0;                                   // 2 x quantised GRU_A state (double-buffered), both halves interleaved (xs_offset)
constexpr uint32_t SM_XB    = SM_XS + 2 * XS_BYTES;                // u32 [2 halves][2 buffers][4 words][16 streams]: quantised GRU_B state
constexpr uint32_t SM_TILES = SM_XB + 2 * 2 * 4 * HALF * 4;
// the tile that held a half's candidate gate is dead once h~ is computed; until the half's next indices are out it carries
constexpr uint32_t T_ACCB   = 0;                                   //   int32 [KPARTS][48][ACCB_ROW] partial sums of the GRU_B input GEMV
constexpr uint32_t T_HBS    = T_ACCB + KPARTS * 3 * NB * ACCB_ROW * 4;   //   float [16 neurons][16 streams] GRU_B state for the sampler
static_assert(T_HBS + NB * HALF * 4 <= TILE_BYTES, "GRU_B scratch must fit inside the gather tile it aliases");
constexpr uint32_t SM_IDX   = SM_TILES + NTILE * TILE_BYTES;       // int32 [2 halves][3][16]: last_sig_ulaw, pred_ulaw, last_exc
constexpr uint32_t SM_MBAR  = al128(SM_IDX + 2 * 3 * HALF * 4);    // mbarriers: image | full[NTILE] | empty[NTILE] | idx[2] | x[2] | accb[2]
constexpr uint32_t MB_IMAGE = SM_MBAR, MB_FULL = SM_MBAR + 8, MB_EMPTY = MB_FULL + 8 * NTILE, MB_IDX = MB_EMPTY + 8 * NTILE, MB_X = MB_IDX + 16, MB_ACCB = MB_X + 16;
constexpr uint32_t SM_IMAGE = SM_MBAR + 128;
static_assert(MB_ACCB + 16 <= SM_IMAGE, "mbarrier block");
// image, fixed part (offsets relative to SM_IMAGE)
constexpr uint32_t SMEM_RESERVED = 1024;                           // shared-window address of dynamic shared memory on sm_100 (checked at kernel start)
constexpr uint32_t IM_LOGIT = 0;                                   // float [256] sampling_logit_table
constexpr uint32_t IM_U2L   = IM_LOGIT + 256 * 4;                  // float [256] ulaw2lin
constexpr uint32_t IM_DIRA  = IM_U2L + 256 * 4;                    // uint32 [NWC][GPW][3][2] = {first quad, quad count}
constexpr uint32_t IM_GRPA  = IM_DIRA + NWC * GPW * 3 * 2 * 4;     // uint32 [NWC][GPW] neuron-group id
constexpr uint32_t IM_DIRB  = IM_GRPA + NWC * GPW * 4;             // uint32 [NWB][2]
constexpr uint32_t IM_PRE_END = IM_DIRB + NWB * 2 * 4;
// u32 [2048] RCPPS table, pre-biased: T[k] + 0x3f800000 (one IADD rebuilds the result); its absolute shared address is
// 8 KB-aligned so that the entry address is table | index (no add)
constexpr uint32_t IM_RCP   = IM_PRE_END + (8192u - (SMEM_RESERVED + SM_IMAGE + IM_PRE_END) % 8192u) % 8192u;
static_assert((SMEM_RESERVED + SM_IMAGE + IM_RCP) % 8192u == 0, "rcp table alignment");
constexpr uint32_t IM_FCWN  = IM_PRE_END;                          // u32: number of dual_fc rows present at IM_FCW (lives in the alignment gap in front of the table)
static_assert(IM_RCP - IM_PRE_END >= 4, "room for the dual_fc row count");
constexpr uint32_t IM_PARA  = IM_RCP + 2048 * 4;                   // float [NWC][GPW][3 gates][16] = recurrent su-bias[8], diag[8]
constexpr uint32_t IM_WBREC = IM_PARA + NWC * GPW * 3 * 16 * 4;    // int8 [6][4][8][4] GRU_B recurrent blocks
constexpr uint32_t IM_PARB  = IM_WBREC + 3 * NB * NB;              // float [96]: input-side su-bias[48], recurrent-side su-bias[48]
constexpr uint32_t IM_FCW   = al128(IM_PARB + 6 * NB * 4);         // float [<= FCW_SMEM_NODES][FCW_ROW] dual_fc rows of the upper tree levels
constexpr uint32_t IM_VAR   = IM_FCW + FCW_SMEM_NODES * FCW_ROW * 4;   // start of the variable-size arrays when all FCW_SMEM_NODES rows are kept

// ---- shared-memory map of the FLOAT-flavour per-sample kernel (sample_kernel_f32.cu) ----
// fp32 GRU_A state tile instead of the u8 one, no gather tiles (per-lane gather), fp16 weights (64 B per block),
// the whole dual_fc table read from global memory.
constexpr int F_NWC = 16, F_GPW = NGRP / F_NWC, F_KPARTS = 2, F_NWB = 6 * F_KPARTS;   // geometry of the float kernel (independent of NWC)
constexpr uint32_t F_XS    = 0;                                    // float [384][32]: GRU_A state of the 32 streams (single buffer)
constexpr uint32_t F_HB    = F_XS + NA * 32 * 4;                   // float [2][16][32]: GRU_B state (double-buffered)
constexpr uint32_t F_ACCB  = F_HB + 2 * NB * 32 * 4;                   // float [48][32]: GRU_B input-side pre-activations
constexpr uint32_t F_IDX   = F_ACCB + 3 * NB * 32 * 4;             // int32 [3][32]
constexpr uint32_t F_MBAR  = al128(F_IDX + 3 * 32 * 4);
constexpr uint32_t F_IMAGE = F_MBAR + 128;
constexpr uint32_t FI_RCP   = 0;                                   // u16 [2048]
constexpr uint32_t FI_LOGIT = FI_RCP + 2048 * 2;
constexpr uint32_t FI_U2L   = FI_LOGIT + 256 * 4;
constexpr uint32_t FI_FCB   = FI_U2L + 256 * 4;
constexpr uint32_t FI_FCF   = FI_FCB + 512 * 4;
constexpr uint32_t FI_PARA  = FI_FCF + 512 * 4;                    // float [F_NWC][F_GPW][3][16] = recurrent bias[8], diag[8]
constexpr uint32_t FI_DIRA  = FI_PARA + F_NWC * F_GPW * 3 * 16 * 4;
constexpr uint32_t FI_GRPA  = FI_DIRA + F_NWC * F_GPW * 3 * 2 * 4;
constexpr uint32_t FI_DIRB  = FI_GRPA + F_NWC * F_GPW * 4;         // uint32 [6*F_KPARTS][2] (only part 0 of each row group is non-empty)
constexpr uint32_t FI_PARB  = FI_DIRB + F_NWB * 2 * 4;               // float [96]: input-side bias[48], recurrent-side bias[48]
constexpr uint32_t FI_VAR   = al128(FI_PARB + 6 * NB * 4);

// ---- FLOAT flavour, small batches: neuron-per-lane kernel (sample_kernel_f32n.cu) ----
// With only a few streams per SM the lane==stream mapping leaves the lanes idle and every lane walks all chains of its
// warp's neurons.  Here a compute lane owns ONE GRU_A neuron (its z, r and h rows: three sequential fp32 FMA chains in the
// reference's block order) for up to FN_S streams of the CTA; the latency of a sample is then set by the longest chain
// (the 384-term rows of GRU_B), not by the work of a whole warp.
constexpr int FN_S = 4;                                             // streams per CTA (live slots: SampleParams::spc <= FN_S)
constexpr int FN_NWC = NA / 32;                                     // 12 compute warps: 384 lanes = 384 neurons
constexpr int FN_THREADS = (FN_NWC + 1) * 32;                       // + sampler warp (lane == stream)
constexpr uint32_t FN_X     = 0;                                    // float [2][FN_S][NA]: GRU_A state (double-buffered)
constexpr uint32_t FN_HB    = FN_X + 2 * FN_S * NA * 4;             // float [2][NB][FN_S]: GRU_B state (double-buffered)
constexpr uint32_t FN_ACCB  = FN_HB + 2 * NB * FN_S * 4;            // float [48][FN_S]: GRU_B input-side pre-activations
constexpr uint32_t FN_IDX   = FN_ACCB + 3 * NB * FN_S * 4;          // int32 [3][FN_S]
constexpr uint32_t FN_MBAR  = al128(FN_IDX + 3 * FN_S * 4);
constexpr uint32_t FN_IMAGE = FN_MBAR + 128;
constexpr uint32_t FNI_RCP   = 0;                                   // u16 [2048]
constexpr uint32_t FNI_LOGIT = FNI_RCP + 2048 * 2;
constexpr uint32_t FNI_U2L   = FNI_LOGIT + 256 * 4;
constexpr uint32_t FNI_FCW   = FNI_U2L + 256 * 4;                   // float [256][FCW_ROW]: all dual_fc rows (weights, biases, factors)
constexpr uint32_t FNI_NEUR  = FNI_FCW + 256 * FCW_ROW * 4;                   // u16 [384]: neuron of compute lane (warp*32 + lane); groups sorted by list length
constexpr uint32_t FNI_DIRA  = FNI_NEUR + NA * 2;                   // u32 [48 groups][3 gates][2] = {first block, block count}
constexpr uint32_t FNI_PARA  = FNI_DIRA + NGRP * 3 * 2 * 4;         // float [3 gates][2][384]: recurrent bias, diag per neuron
constexpr uint32_t FNI_DIRB  = FNI_PARA + 3 * 2 * NA * 4;           // u32 [6][2]
constexpr uint32_t FNI_PARB  = FNI_DIRB + 64;                       // float [96]
constexpr uint32_t FNI_WBREC = FNI_PARB + 6 * NB * 4;               // float [16 in][48 out]
constexpr uint32_t FNI_VAR   = al128(FNI_WBREC + 3 * NB * NB * 4);  // blocks: fp16 [8 rows][4 cols] (64 B), u16 meta = 4 * pos

// offsets used by the (flavour-agnostic) image builder
struct ImageMap { uint32_t sm_image, rcp, logit, u2l, fcw, fcb, fcf, parA, dirA, grpA, dirB, wBrec, parB, var; };
constexpr ImageMap MAP_INT8 = {SM_IMAGE, IM_RCP, IM_LOGIT, IM_U2L, IM_FCW, 0xFFFFFFFFu, 0xFFFFFFFFu, IM_PARA, IM_DIRA, IM_GRPA, IM_DIRB, IM_WBREC, IM_PARB, IM_VAR};
constexpr ImageMap MAP_F32  = {F_IMAGE, FI_RCP, FI_LOGIT, FI_U2L, 0xFFFFFFFFu, FI_FCB, FI_FCF, FI_PARA, FI_DIRA, FI_GRPA, FI_DIRB, 0xFFFFFFFFu, FI_PARB, FI_VAR};

struct SmemLayout {          // run-time part; offsets are absolute (from the start of dynamic shared memory)
    uint32_t wA;        // GRU_A weights, ordered (warp, slot, gate): int8 flavour 128-byte quads, float flavour 64-byte fp16 blocks
    uint32_t metaA;     // int8: 4 x u16 per quad (xs_offset of each slot's column block); float: u16 per block
    uint32_t wB;        // GRU_B input weights, ordered (row group, K part)
    uint32_t metaB;
    uint32_t wBrecF;    // float flavour, lane==stream image: fp32 [16 in][48 out] GRU_B recurrent weights.  Neuron-per-lane image (its
                        // recurrent weights sit at FNI_WBREC): 1 if every GRU_B row group lists all 96 column blocks in order
    uint32_t sm_image;  // where the image starts (SM_IMAGE for int8, F_IMAGE for the float flavour)
    uint32_t image_bytes;
    uint32_t total_bytes;
    uint32_t nblkA_padded, nblkB_padded;   // units in wA / wB: quads (int8 flavour) or blocks (float flavour)
};

// Device-resident model (one per batch; weights replicated per GPU, ~4 MB).
struct DeviceModel {
    int is_float;                    // 0: int8 DOT_PROD semantics (oracle A); 1: float semantics (oracle B)
    int fast_cvt;                    // int8: GRU_A pre-activations provably < 2^22 accumulator units (conversion-free rounding allowed)
    float lpc_gamma;
    SmemLayout L;
    uint8_t *image;                  // [L.image_bytes] global copy of the SMEM image
    SmemLayout Ln;                   // float flavour: layout / image of the neuron-per-lane kernel (small batches)
    uint8_t *image_n;
    // per-sample gathers (L2-resident): [256][3*NA] each
    float *emb_sig, *emb_pred, *emb_exc;
    float *fcw;                      // dual_fc rows [256][FCW_ROW] = 32 weights, 2 biases, 2 factors (the lower tree levels are read from here)
    // frame network (fp32, reference layouts kept: column-major W[j*N+i], conv W[(k*in+i)*out+o])
    float *embed_pitch;              // [256][64]
    float *conv1_w, *conv1_b, *conv2_w, *conv2_b;
    float *dense1_w, *dense1_b, *dense2_w, *dense2_b;
    float *gad_w, *gad_b, *gbd_w, *gbd_b;
    // tables
    uint16_t *rcp16;                 // [2048]
    float *dct;                      // [18*18]
    float *twiddles;                 // [320*2]
    int16_t *bitrev;                 // [320]
    float *gamma_pow;                // [16] gamma^(i+1) accumulated in float exactly like lpc_weighting (freq.c:299-308)
    float *pitch_pow;                // [64] (float)(pow(2.f, k/21.)*32) for the packet decoder (lpcnet_dec.c:124)
    float *codebooks;                // cb1,cb2,cb3,diff4 or NULL
    // accounting
    long algo_bytes_total, algo_bytes_sparse;
    int nblkA, nblkB;
};

struct SampleParams {
    SmemLayout L;
    const uint8_t *image;
    const float *emb_sig, *emb_pred, *emb_exc;
    const float *fcw;        // [256][FCW_ROW] dual_fc rows (levels 6,7 of the sampling tree)
    const float *condA;      // [nframes][n][3*NA]
    const float *condB;      // [nframes][n][3*NB]
    const float *lpc_raw;    // [nframes + 2][n][16]  (frame f uses entry f: the LPC of frame f-2)
    const float *gamma_pow;  // [16]
    float *hA;               // [NA][n]
    float *hB;               // [NB][n]
    float *last_sig;         // [16][n]
    float *deemph;           // [n]
    int *last_exc;           // [n]
    uint32_t *rng;           // [4][n]
    short *pcm;              // stream s, frame f, sample t at pcm[s*pcm_stream_stride + f*spf + t]
    long long pcm_stream_stride;
    int n_streams, nframes, spf;
    int fast_cvt;
    int spc;                 // live streams per CTA (1..32): chosen by the launcher so that the grid covers the SMs
    int one_half;            // spc <= 16: all live streams sit in half A and half B is not stepped at all (a small batch is bound by the
                             // latency of one sample, which the second half would only lengthen)
#ifdef LPCNET_TRACE
    long long *trace;        // tuning builds only: clock64 stamps of CTA 0, [8 samples][32 events]
#endif
};

// ---- host-side API of the internal modules ----
int model_load(DeviceModel *m, const unsigned char *blob, int len, float lpc_gamma);   // 0 / -1 (sets error)
void model_free(DeviceModel *m);
int debug_build_image(const unsigned char *blob, int len, unsigned char *out, size_t cap, SmemLayout *L);
int debug_build_image_n(const unsigned char *blob, int len, unsigned char *out, size_t cap, SmemLayout *L);
void set_error(const char *fmt, ...);

struct FrameState {           // per-batch persistent state of the 100 Hz path
    float *conv1_state;       // [n][FRAME_IN*2]
    float *conv2_state;       // [n][COND*2]
    float *lpc_carry;         // [2][n][16] raw LPC of the two previous frames
    float *vq_mem;            // [n][18] decoder memory
};

void launch_frame_network(const DeviceModel &m, const FrameState &fs, const float *d_features, long long stream_stride,
                          int frame_stride, int n, int nframes, int frame_count0, float *condA, float *condB,
                          float *lpc_raw /* [nframes+2][n][16], entries 0,1 = carry */, cudaStream_t st);
void launch_decode_packets(const DeviceModel &m, const FrameState &fs, const uint8_t *d_packets, int n, int npackets,
                           float *d_features /* [n][4*npackets][20] */, cudaStream_t st);
cudaError_t launch_sample_kernel(const SampleParams &p, cudaStream_t st);
cudaError_t launch_sample_kernel_f32(const SampleParams &p, cudaStream_t st);
cudaError_t launch_sample_kernel_f32n(const SampleParams &p, cudaStream_t st);   // p.L / p.image = DeviceModel::Ln / image_n, p.spc <= FN_S
int sample_kernel_smem_ok(uint32_t bytes);
int streams_per_cta_for(int n_streams);   // min(32, ceil(n / SM count)) on the current device

}  // namespace lpcnet_b200
