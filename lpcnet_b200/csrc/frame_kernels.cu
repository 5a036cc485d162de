// frame_kernels.cu — the 100 Hz path: conditioning network, cepstrum->LPC, packet decoder.
//
// Replaces (reference file:line):
//   run_frame_network          src/lpcnet.c:82-120      (embedding, conv1d x2, dense x4 -> gru_a/gru_b condition)
//   compute_conv1d / dense     src/nnet.c:452-470,122-135 over sgemv_accum16 (src/vec_avx.h:618-643)
//   lpc_from_cepstrum          src/freq.c:310-320 (+ idct :230, interp_band_gain :202, inverse_transform :256,
//                              opus_fft_c src/kiss_fft.c:566, lpcn_lpc src/freq.c:86)
//   decode_packet              src/lpcnet_dec.c:81-155, perform_double_interp src/common.c:58-65
//
// Arithmetic contract: every output neuron is ONE thread walking its inputs in ascending order with __fmaf_rn
// (exactly the per-row FMA chain of sgemv_accum16), so results equal the reference build A/B bit for bit.  This
// file is compiled with -fmad=false: plain a*b+c below are separate IEEE operations, as in the pinned oracle build
// (-ffp-contract=off).  ~2.5 % of the MACs of the whole path: not on the roofline-critical side.
#include <cstdint>
#include "engine.h"
#include "devmath.cuh"

namespace lpcnet_b200 {

constexpr int TS = 8;                  // streams per block (register tile over streams: one weight load feeds 8 FMAs)
constexpr int FT = 128;                // threads per block == COND

// one dense/conv layer for TS streams: out[i][k] = act(b[i] + sum_j W[j*N+i]*x[j][k]); thread i = threadIdx.x + o*FT
template <int M>
__device__ __forceinline__ void layer_accum(float y[TS], const float *__restrict__ W, int N, int i, const float (*x)[TS])
{
#pragma unroll 4
    for (int j = 0; j < M; j++) {
        const float w = __ldg(&W[(size_t)j * N + i]);
        const float4 a = *reinterpret_cast<const float4 *>(&x[j][0]);
        const float4 b = *reinterpret_cast<const float4 *>(&x[j][4]);
        y[0] = __fmaf_rn(w, a.x, y[0]); y[1] = __fmaf_rn(w, a.y, y[1]); y[2] = __fmaf_rn(w, a.z, y[2]); y[3] = __fmaf_rn(w, a.w, y[3]);
        y[4] = __fmaf_rn(w, b.x, y[4]); y[5] = __fmaf_rn(w, b.y, y[5]); y[6] = __fmaf_rn(w, b.z, y[6]); y[7] = __fmaf_rn(w, b.w, y[7]);
    }
}

struct FrameNetArgs {
    const float *embed_pitch, *conv1_w, *conv1_b, *conv2_w, *conv2_b, *dense1_w, *dense1_b, *dense2_w, *dense2_b;
    const float *gad_w, *gad_b, *gbd_w, *gbd_b;
    const uint16_t *rcp16;
    float *conv1_state, *conv2_state;
    const float *features; long long stream_stride; int frame_stride;
    int n, nframes;
    int *frame_count;        // [n] per-stream frame counter (lpcnet.c:119), read at entry, advanced by nframes (saturating at 1000) at exit
    int na;                  // GRU_A units: gru_a_dense_feature has 3*na outputs
    int features_delay;      // FEATURES_DELAY of the model (conv2 warm-up zeroing, lpcnet.c:101)
    float *condA, *condB;
    float *lpc_e2e;          // END2END models: [nframes][n][16] LPC from the network's reflection coefficients (lpcnet.c:107-108), else NULL
};

// rc2lpc (lpcnet.c:57-78): reflection coefficients -> direct-form LPC, the reference's exact (unusual) recursion
__device__ __forceinline__ void rc2lpc_dev(float *lpc, const float *rc)
{
    float tmp[LPC_ORDER], ntmp[LPC_ORDER];
    for (int i = 0; i < LPC_ORDER; i++) { tmp[i] = rc[i]; ntmp[i] = 0.f; }
    for (int i = 0; i < LPC_ORDER; i++) {
        for (int j = 0; j <= i - 1; j++) ntmp[j] = __fadd_rn(tmp[j], __fmul_rn(tmp[i], tmp[i - j - 1]));
        for (int k = 0; k <= i - 1; k++) tmp[k] = ntmp[k];
    }
    for (int i = 0; i < LPC_ORDER; i++) lpc[i] = tmp[i];
}

__global__ void __launch_bounds__(FT) frame_net_kernel(const FrameNetArgs a)
{
    __shared__ __align__(16) float xin[3 * FRAME_IN][TS];   // conv1 input window: [2 old frames | current]
    __shared__ __align__(16) float x2[3 * COND][TS];        // conv2 input window
    __shared__ __align__(16) float c2[COND][TS], d1[COND][TS], cd[COND][TS];
    __shared__ int pitch_s[TS];
    __shared__ int fc_s[TS];                                // frame_count of the block's streams
    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * TS;
    auto sid = [&](int k) { return min(s0 + k, a.n - 1); };   // tail block: replicate the last stream, stores masked

    for (int e = tid; e < 2 * FRAME_IN * TS; e += FT) { int j = e / TS, k = e % TS; xin[j][k] = a.conv1_state[(size_t)sid(k) * 2 * FRAME_IN + j]; }
    for (int e = tid; e < 2 * COND * TS; e += FT) { int j = e / TS, k = e % TS; x2[j][k] = a.conv2_state[(size_t)sid(k) * 2 * COND + j]; }
    if (tid < TS) fc_s[tid] = a.frame_count[sid(tid)];
    __syncthreads();

    for (int f = 0; f < a.nframes; f++) {
        // ---- input assembly (lpcnet.c:93-96) ----
        if (tid < TS) {
            const float *ft = a.features + (size_t)sid(tid) * a.stream_stride + (size_t)f * a.frame_stride;
            // pitch = (int)floor(.1 + 50*features[NB_BANDS]+100): 50*f is a float product, the sums are double
            int p = (int)floor((.1 + (double)__fmul_rn(50.f, ft[NB_BANDS])) + 100.0);
            pitch_s[tid] = min(255, max(33, p));
        }
        for (int e = tid; e < NB_FEAT * TS; e += FT) {
            int j = e / TS, k = e % TS;
            xin[2 * FRAME_IN + j][k] = a.features[(size_t)sid(k) * a.stream_stride + (size_t)f * a.frame_stride + j];
        }
        __syncthreads();
        for (int e = tid; e < PITCH_EMBED * TS; e += FT) {
            int j = e / TS, k = e % TS;
            xin[2 * FRAME_IN + NB_FEAT + j][k] = __ldg(&a.embed_pitch[pitch_s[k] * PITCH_EMBED + j]);
        }
        __syncthreads();
        float y[TS];
        // ---- conv1 (nnet.c:452-470; zeroed while frame_count < FEATURE_CONV1_DELAY=1, lpcnet.c:99) ----
        {
            const float b = __ldg(&a.conv1_b[tid]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<3 * FRAME_IN>(y, a.conv1_w, COND, tid, xin);
#pragma unroll
            for (int k = 0; k < TS; k++) x2[2 * COND + tid][k] = fc_s[k] < 1 ? 0.f : tanh_approx(y[k], a.rcp16);
        }
        __syncthreads();
        // conv1 window shift, mem <- tmp[nb_inputs:] (nnet.c:469): two passes through registers because source and
        // destination rows overlap
        {
            float tmpv[(2 * FRAME_IN * TS + FT - 1) / FT];
            int c = 0;
            for (int e = tid; e < 2 * FRAME_IN * TS; e += FT, c++) tmpv[c] = xin[FRAME_IN + e / TS][e % TS];
            __syncthreads();
            c = 0;
            for (int e = tid; e < 2 * FRAME_IN * TS; e += FT, c++) xin[e / TS][e % TS] = tmpv[c];
        }
        // ---- conv2 (zeroed while frame_count < FEATURES_DELAY, lpcnet.c:101) ----
        {
            const float b = __ldg(&a.conv2_b[tid]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<3 * COND>(y, a.conv2_w, COND, tid, x2);
#pragma unroll
            for (int k = 0; k < TS; k++) c2[tid][k] = fc_s[k] < a.features_delay ? 0.f : tanh_approx(y[k], a.rcp16);
        }
        __syncthreads();
        {
            float tmpv[(2 * COND * TS + FT - 1) / FT];
            int c = 0;
            for (int e = tid; e < 2 * COND * TS; e += FT, c++) tmpv[c] = x2[COND + e / TS][e % TS];
            __syncthreads();
            c = 0;
            for (int e = tid; e < 2 * COND * TS; e += FT, c++) x2[e / TS][e % TS] = tmpv[c];
        }
        // ---- dense1, dense2 (tanh) ----
        {
            const float b = __ldg(&a.dense1_b[tid]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<COND>(y, a.dense1_w, COND, tid, c2);
#pragma unroll
            for (int k = 0; k < TS; k++) d1[tid][k] = tanh_approx(y[k], a.rcp16);
        }
        __syncthreads();
        {
            const float b = __ldg(&a.dense2_b[tid]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<COND>(y, a.dense2_w, COND, tid, d1);
#pragma unroll
            for (int k = 0; k < TS; k++) cd[tid][k] = tanh_approx(y[k], a.rcp16);
        }
        __syncthreads();
        // ---- END2END: the first 16 conditioning outputs are reflection coefficients (lpcnet.c:105,107-108) ----
        if (a.lpc_e2e && tid < TS && s0 + tid < a.n) {
            float rc[LPC_ORDER], lp[LPC_ORDER];
            for (int i = 0; i < LPC_ORDER; i++) rc[i] = cd[i][tid];
            rc2lpc_dev(lp, rc);
            float *o = a.lpc_e2e + ((size_t)f * a.n + s0 + tid) * LPC_ORDER;
            for (int i = 0; i < LPC_ORDER; i++) o[i] = lp[i];
        }
        // ---- gru_a_dense_feature (128 -> 3*na, linear) and gru_b_dense_feature (128 -> 48, linear) ----
        const int na3 = 3 * a.na;
        for (int o = 0; o < na3 / FT; o++) {
            const int i = o * FT + tid;
            const float b = __ldg(&a.gad_b[i]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<COND>(y, a.gad_w, na3, i, cd);
#pragma unroll
            for (int k = 0; k < TS; k++) if (s0 + k < a.n) a.condA[((size_t)f * a.n + s0 + k) * na3 + i] = y[k];
        }
        if (tid < 3 * NB) {
            const float b = __ldg(&a.gbd_b[tid]);
#pragma unroll
            for (int k = 0; k < TS; k++) y[k] = b;
            layer_accum<COND>(y, a.gbd_w, 3 * NB, tid, cd);
#pragma unroll
            for (int k = 0; k < TS; k++) if (s0 + k < a.n) a.condB[((size_t)f * a.n + s0 + k) * (3 * NB) + tid] = y[k];
        }
        __syncthreads();
        if (tid < TS && fc_s[tid] < 1000) fc_s[tid]++;      // lpcnet.c:119 (visible to the next frame after its first barrier)
    }
    __syncthreads();
    if (tid < TS && s0 + tid < a.n) a.frame_count[s0 + tid] = fc_s[tid];
    for (int e = tid; e < 2 * FRAME_IN * TS; e += FT) { int j = e / TS, k = e % TS; if (s0 + k < a.n) a.conv1_state[(size_t)(s0 + k) * 2 * FRAME_IN + j] = xin[j][k]; }
    for (int e = tid; e < 2 * COND * TS; e += FT) { int j = e / TS, k = e % TS; if (s0 + k < a.n) a.conv2_state[(size_t)(s0 + k) * 2 * COND + j] = x2[j][k]; }
}

// ------------------------------------------------------------------------------------------------------------------
// cepstrum -> LPC, one thread per (frame, stream).  The 320-point complex FFT follows the reference's mixed-radix
// schedule (factors 5,4,4,4: lpcnet_tables.c:200) butterfly for butterfly so that the 17 autocorrelation lags — and
// hence the LPCs, the prediction and every u-law index derived from it — are bit-identical.
struct c32 { float r, i; };
#define CMUL(m_, a_, b_) do { (m_).r = (a_).r * (b_).r - (a_).i * (b_).i; (m_).i = (a_).r * (b_).i + (a_).i * (b_).r; } while (0)
#define CADD(r_, a_, b_) do { (r_).r = (a_).r + (b_).r; (r_).i = (a_).i + (b_).i; } while (0)
#define CSUB(r_, a_, b_) do { (r_).r = (a_).r - (b_).r; (r_).i = (a_).i - (b_).i; } while (0)

__device__ void radix4(c32 *Fout, int fstride, const c32 *__restrict__ tw, int m, int N, int mm)   // kiss_fft.c:101-170
{
    if (m == 1) {
        for (int i = 0; i < N; i++) {
            c32 s0, s1;
            CSUB(s0, Fout[0], Fout[2]); CADD(Fout[0], Fout[0], Fout[2]);
            CADD(s1, Fout[1], Fout[3]); CSUB(Fout[2], Fout[0], s1); CADD(Fout[0], Fout[0], s1);
            CSUB(s1, Fout[1], Fout[3]);
            Fout[1].r = s0.r + s1.i; Fout[1].i = s0.i - s1.r;
            Fout[3].r = s0.r - s1.i; Fout[3].i = s0.i + s1.r;
            Fout += 4;
        }
    } else {
        c32 *beg = Fout; const int m2 = 2 * m, m3 = 3 * m;
        for (int i = 0; i < N; i++) {
            Fout = beg + i * mm;
            for (int j = 0; j < m; j++) {
                c32 s0, s1, s2, s3, s4, s5;
                const c32 t1 = tw[j * fstride], t2 = tw[j * fstride * 2], t3 = tw[j * fstride * 3];
                CMUL(s0, Fout[m], t1); CMUL(s1, Fout[m2], t2); CMUL(s2, Fout[m3], t3);
                CSUB(s5, Fout[0], s1); CADD(Fout[0], Fout[0], s1);
                CADD(s3, s0, s2); CSUB(s4, s0, s2);
                CSUB(Fout[m2], Fout[0], s3);
                CADD(Fout[0], Fout[0], s3);
                Fout[m].r = s5.r + s4.i; Fout[m].i = s5.i - s4.r;
                Fout[m3].r = s5.r - s4.i; Fout[m3].i = s5.i + s4.r;
                ++Fout;
            }
        }
    }
}
__device__ void radix5_last(c32 *F0, const c32 *__restrict__ tw)     // kiss_fft.c:232-311 with m=64, N=1, fstride=1
{
    const int m = 64;
    const c32 ya = tw[m], yb = tw[2 * m];
    c32 *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
    for (int u = 0; u < m; ++u) {
        c32 s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12;
        s0 = *F0;
        CMUL(s1, *F1, tw[u]); CMUL(s2, *F2, tw[2 * u]); CMUL(s3, *F3, tw[3 * u]); CMUL(s4, *F4, tw[4 * u]);
        CADD(s7, s1, s4); CSUB(s10, s1, s4); CADD(s8, s2, s3); CSUB(s9, s2, s3);
        F0->r = F0->r + (s7.r + s8.r); F0->i = F0->i + (s7.i + s8.i);
        s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r); s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
        s6.r = s10.i * ya.i + s9.i * yb.i; s6.i = -(s10.r * ya.i + s9.r * yb.i);
        CSUB(*F1, s5, s6); CADD(*F4, s5, s6);
        s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r); s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
        s12.r = s9.i * ya.i - s10.i * yb.i; s12.i = s10.r * yb.i - s9.r * ya.i;
        CADD(*F2, s11, s12); CSUB(*F3, s11, s12);
        ++F0; ++F1; ++F2; ++F3; ++F4;
    }
}

struct LpcArgs {
    const float *features; long long stream_stride; int frame_stride; int n, nframes;
    const float *dct; const c32 *tw; const int16_t *bitrev;
    float *lpc_raw;       // [nframes+2][n][16]; this kernel fills entries 2..nframes+1
};

__global__ void __launch_bounds__(64) lpc_kernel(const LpcArgs a)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.n * a.nframes) return;
    const int f = (int)(gid / a.n), s = (int)(gid % a.n);
    const float *cep = a.features + (size_t)s * a.stream_stride + (size_t)f * a.frame_stride;
    const short eband5ms[NB_BANDS] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40};               // freq.c:45-48
    const float compensation[NB_BANDS] = {0.8f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.666667f, 0.5f, 0.5f, 0.5f,
                                          0.333333f, 0.25f, 0.25f, 0.2f, 0.166667f, 0.173913f};                  // freq.c:50-52
    float tmp[NB_BANDS], Ex[NB_BANDS];
    for (int i = 0; i < NB_BANDS; i++) tmp[i] = cep[i];
    tmp[0] += 4;
    const double idct_scale = sqrt(2. / NB_BANDS);
    for (int i = 0; i < NB_BANDS; i++) {                          // idct freq.c:230-240
        float sum = 0;
        for (int j = 0; j < NB_BANDS; j++) sum += tmp[j] * __ldg(&a.dct[i * NB_BANDS + j]);
        Ex[i] = (float)((double)sum * idct_scale);
    }
    for (int i = 0; i < NB_BANDS; i++) Ex[i] = (float)(pow(10.0, (double)Ex[i]) * (double)compensation[i]);   // freq.c:318
    // interp_band_gain (freq.c:202-215) + Hermitian extension (freq.c:256-266) + scale & digit-reverse (kiss_fft.c:575-584)
    c32 y[WINDOW_SIZE];
    {
        float Xr[FREQ_SIZE];
        for (int i = 0; i < FREQ_SIZE; i++) Xr[i] = 0;
        for (int i = 0; i < NB_BANDS - 1; i++) {
            int band_size = (eband5ms[i + 1] - eband5ms[i]) * 4;
            for (int j = 0; j < band_size; j++) {
                float frac = __fdiv_rn((float)j, (float)band_size);
                Xr[(eband5ms[i] * 4) + j] = (1 - frac) * Ex[i] + frac * Ex[i + 1];
            }
        }
        Xr[FREQ_SIZE - 1] = 0;
        const float scale = 1.f / WINDOW_SIZE;
        for (int i = 0; i < WINDOW_SIZE; i++) {
            float xr = i < FREQ_SIZE ? Xr[i] : Xr[WINDOW_SIZE - i];
            float xi = i < FREQ_SIZE ? 0.f : -0.f;
            int o = a.bitrev[i];
            y[o].r = scale * xr; y[o].i = scale * xi;
        }
    }
    radix4(y, 80, a.tw, 1, 80, 4);
    radix4(y, 20, a.tw, 4, 20, 16);
    radix4(y, 5, a.tw, 16, 5, 64);
    radix5_last(y, a.tw);
    float ac[LPC_ORDER + 1];
    ac[0] = WINDOW_SIZE * y[0].r;
    for (int i = 1; i < LPC_ORDER + 1; i++) ac[i] = WINDOW_SIZE * y[WINDOW_SIZE - i].r;
    ac[0] = (float)((double)ac[0] + ((double)ac[0] * 1e-4 + 320 / 12 / 38.));       // freq.c:292
    for (int i = 1; i < LPC_ORDER + 1; i++) ac[i] = (float)((double)ac[i] * (1 - 6e-5 * i * i));   // freq.c:294
    float lpc[LPC_ORDER];
    for (int i = 0; i < LPC_ORDER; i++) lpc[i] = 0;
    {                                                             // lpcn_lpc freq.c:86-127 (float build)
        float error = ac[0];
        if (ac[0] != 0) {
            for (int i = 0; i < LPC_ORDER; i++) {
                float rr = 0;
                for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
                rr += ac[i + 1];
                float r = __fdiv_rn(-rr, error);
                lpc[i] = r;
                for (int j = 0; j < (i + 1) >> 1; j++) {
                    float t1 = lpc[j], t2 = lpc[i - 1 - j];
                    lpc[j] = t1 + r * t2;
                    lpc[i - 1 - j] = t2 + r * t1;
                }
                error = error - (r * r) * error;
                if (error < .001f * ac[0]) break;
            }
        }
    }
    float *out = a.lpc_raw + ((size_t)(f + 2) * a.n + s) * LPC_ORDER;
    for (int i = 0; i < LPC_ORDER; i++) out[i] = lpc[i];
}

// lpc_carry[0] = LPC of frame -1, lpc_carry[1] = frame -2 (old_lpc[0], old_lpc[1] of lpcnet.c:110-112).
// lpc_raw entry e holds the raw LPC of frame e-2, so frame f reads entry f.
__global__ void lpc_carry_in_kernel(const float *carry, float *lpc_raw, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * LPC_ORDER) return;
    lpc_raw[i] = carry[(size_t)n * LPC_ORDER + i];                       // entry 0 = frame -2
    lpc_raw[(size_t)n * LPC_ORDER + i] = carry[i];                       // entry 1 = frame -1
}
__global__ void lpc_carry_out_kernel(float *carry, const float *lpc_raw, int n, int nframes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * LPC_ORDER) return;
    carry[i] = lpc_raw[(size_t)(nframes + 1) * n * LPC_ORDER + i];              // last frame
    carry[(size_t)n * LPC_ORDER + i] = lpc_raw[(size_t)nframes * n * LPC_ORDER + i];   // the one before
}

// lpc_raw [nframes+2][n][16]: entry e holds the raw (unweighted) LPC of frame e-2 of this call (entries 0,1 = the carry of the
// previous call), so a model with FEATURES_DELAY d reads entry f + 2 - d for frame f (lpcnet.c:109-115); END2END models get the
// network's own LPC of frame f written to entry f + 2 (and are read with d = 0).
void launch_frame_network(const DeviceModel &m, const FrameState &fs, const float *d_features, long long stream_stride,
                          int frame_stride, int n, int nframes, float *condA, float *condB, float *lpc_raw, cudaStream_t st)
{
    FrameNetArgs fa{m.embed_pitch, m.conv1_w, m.conv1_b, m.conv2_w, m.conv2_b, m.dense1_w, m.dense1_b, m.dense2_w, m.dense2_b,
                    m.gad_w, m.gad_b, m.gbd_w, m.gbd_b, m.rcp16, fs.conv1_state, fs.conv2_state, d_features, stream_stride,
                    frame_stride, n, nframes, fs.frame_count, m.na, m.cfg.features_delay, condA, condB,
                    m.cfg.end2end ? lpc_raw + (size_t)2 * n * LPC_ORDER : nullptr};
    frame_net_kernel<<<(n + TS - 1) / TS, FT, 0, st>>>(fa);
    if (m.cfg.end2end) return;                                 // no cepstrum -> LPC path, no delay line (lpcnet.c:107-108)
    const int tpb = 128;
    lpc_carry_in_kernel<<<(n * LPC_ORDER + tpb - 1) / tpb, tpb, 0, st>>>(fs.lpc_carry, lpc_raw, n);
    LpcArgs la{d_features, stream_stride, frame_stride, n, nframes, m.dct, reinterpret_cast<const c32 *>(m.twiddles), m.bitrev, lpc_raw};
    long long tot = (long long)n * nframes;
    lpc_kernel<<<(unsigned)((tot + 63) / 64), 64, 0, st>>>(la);
    lpc_carry_out_kernel<<<(n * LPC_ORDER + tpb - 1) / tpb, tpb, 0, st>>>(fs.lpc_carry, lpc_raw, n, nframes);
}

// ------------------------------------------------------------------------------------------------------------------
// decode_packet (lpcnet_dec.c:81-155): 64-bit packet -> 4 feature frames; one thread per stream, packets in order
// (vq_mem carries across packets).
struct DecodeArgs { const uint8_t *packets; int n, npackets; const float *cb; const float *pitch_pow; float *vq_mem; float *features; };

__global__ void decode_kernel(const DecodeArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const float *cb1 = a.cb, *cb2 = cb1 + 1024 * 17, *cb3 = cb2 + 1024 * 17, *cbd4 = cb3 + 1024 * 17;
    float vq_mem[NB_BANDS];
    for (int i = 0; i < NB_BANDS; i++) vq_mem[i] = a.vq_mem[(size_t)s * NB_BANDS + i];
    for (int p = 0; p < a.npackets; p++) {
        const uint8_t *buf = a.packets + ((size_t)s * a.npackets + p) * 8;
        unsigned long long bits = 0;
        for (int i = 0; i < 8; i++) bits = (bits << 8) | buf[i];                  // MSB-first, bits_unpack :59-78
        int pos = 0;
        auto get = [&](int nb) { pos += nb; return (int)((bits >> (64 - pos)) & ((1ull << nb) - 1)); };
        int c0_id = get(7), main_pitch = get(6), modulation = get(3), corr_id = get(2);
        int vq0 = get(10), vq1 = get(10), vq2 = get(10), vq_mid = get(13), interp_id = get(3);
        float *F = a.features + ((size_t)s * a.npackets + p) * 4 * NB_FEAT;      // [4][20]
        float f1[NB_BANDS], f3[NB_BANDS];
        int voiced = 1;
        modulation -= 4;
        if (modulation == -4) { voiced = 0; modulation = 0; }
        float frame_corr = voiced ? 0.3875f + .175f * corr_id : 0.0375f + .075f * corr_id;
        for (int sub = 0; sub < 4; sub++) {
            float pch = a.pitch_pow[main_pitch];                                  // (float)(pow(2.f, main_pitch/21.)*PITCH_MIN_PERIOD)
            pch *= 1.f + __fdiv_rn(__fdiv_rn((float)modulation, 16.f), 7.f) * (2 * sub - 3);
            float mx = 33 > pch ? 33.f : pch;                                     // MIN16(255, MAX16(33, p))
            pch = 255 < mx ? 255.f : mx;
            F[sub * NB_FEAT + NB_BANDS] = .02f * (pch - 100.f);
            F[sub * NB_FEAT + NB_BANDS + 1] = frame_corr - .5f;
        }
        f3[0] = __fdiv_rn((float)(c0_id - 64), 4.f);
        for (int i = 0; i < NB_BANDS - 1; i++) f3[i + 1] = cb1[vq0 * 17 + i] + cb2[vq1 * 17 + i] + cb3[vq2 * 17 + i];
        float sign = 1;
        if (vq_mid >= 4096) { vq_mid -= 4096; sign = -1; }
        for (int i = 0; i < NB_BANDS; i++) f1[i] = sign * cbd4[vq_mid * NB_BANDS + i];
        if ((vq_mid & 3) < 2) { for (int i = 0; i < NB_BANDS; i++) f1[i] += .5f * (vq_mem[i] + f3[i]); }
        else if ((vq_mid & 3) == 2) { for (int i = 0; i < NB_BANDS; i++) f1[i] += vq_mem[i]; }
        else { for (int i = 0; i < NB_BANDS; i++) f1[i] += f3[i]; }
        interp_id += (interp_id >= 7);                                            // FORBIDDEN_INTERP, common.c:58-65
        const int id0 = interp_id / 3, id1 = interp_id % 3;
        for (int i = 0; i < NB_BANDS; i++) {
            F[0 * NB_FEAT + i] = id0 == 0 ? .5f * (vq_mem[i] + f1[i]) : (id0 == 1 ? vq_mem[i] : f1[i]);
            F[1 * NB_FEAT + i] = f1[i];
            F[2 * NB_FEAT + i] = id1 == 0 ? .5f * (f1[i] + f3[i]) : (id1 == 1 ? f1[i] : f3[i]);
            F[3 * NB_FEAT + i] = f3[i];
            vq_mem[i] = f3[i];
        }
    }
    for (int i = 0; i < NB_BANDS; i++) a.vq_mem[(size_t)s * NB_BANDS + i] = vq_mem[i];
}

void launch_decode_packets(const DeviceModel &m, const FrameState &fs, const uint8_t *d_packets, int n, int npackets,
                           float *d_features, cudaStream_t st)
{
    DecodeArgs da{d_packets, n, npackets, m.codebooks, m.pitch_pow, fs.vq_mem, d_features};
    decode_kernel<<<(n + 127) / 128, 128, 0, st>>>(da);
}

}  // namespace lpcnet_b200
