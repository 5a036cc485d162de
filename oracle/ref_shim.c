/* TEST INFRASTRUCTURE — compiled together with the untouched reference sources (oracle/Makefile `ref`).
 * Gives the Python tests a few convenience entry points around the reference's own public API
 * (include/lpcnet.h) and exposes static-inline helpers of the reference (common.h, vec_avx.h) by value.
 * Nothing here re-implements reference behaviour. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <immintrin.h>
#include "nnet_data.h"
#include "nnet.h"
#include "common.h"
#include "freq.h"
#include "lpcnet.h"
#include "lpcnet_private.h"

/* _mm_rcp_ss on mantissa bin k (top 11 mantissa bits), exponent 0: the table the engine must emulate. */
void ref_rcp_table(uint32_t *out)
{
    int k;
    for (k = 0; k < 2048; k++) {
        union { float f; uint32_t u; } in, o;
        in.u = 0x3f800000u | ((uint32_t)k << 12);
        o.f = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(in.f)));
        out[k] = o.u;
    }
}

float ref_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
int ref_lin2ulaw(float x) { return lin2ulaw(x); }
float ref_ulaw2lin(float u) { return ulaw2lin(u); }
void ref_activation(float *out, const float *in, int n, int act) { compute_activation(out, in, n, act); }
int ref_state_size(void) { return lpcnet_get_size(); }
int ref_is_float_build(void)
{
#ifdef DOT_PROD
    return 0;
#else
    return 1;
#endif
}

/* Synthesise nframes frames (features: [nframes][stride] floats, first 20 used) with a fresh state. */
int ref_synth_stream(const unsigned char *blob, int len, const float *features, int stride, int nframes,
                     short *pcm)
{
    int i;
    LPCNetState *st = lpcnet_create();
    if (lpcnet_load_model(st, blob, len) != 0) { lpcnet_destroy(st); return -1; }
    for (i = 0; i < nframes; i++)
        lpcnet_synthesize(st, features + (size_t)i * stride, pcm + (size_t)i * LPCNET_FRAME_SIZE, LPCNET_FRAME_SIZE);
    lpcnet_destroy(st);
    return 0;
}

/* Decode npackets 8-byte packets with a fresh decoder state. LPCNetDecState starts with its LPCNetState
 * (lpcnet_private.h:50-53) so the blob is loaded through that member. */
int ref_decode_stream(const unsigned char *blob, int len, const unsigned char *packets, int npackets, short *pcm)
{
    int i;
    LPCNetDecState *st = lpcnet_decoder_create();
    if (lpcnet_load_model((LPCNetState *)st, blob, len) != 0) { lpcnet_decoder_destroy(st); return -1; }
    for (i = 0; i < npackets; i++)
        lpcnet_decode(st, packets + (size_t)i * LPCNET_COMPRESSED_SIZE, pcm + (size_t)i * LPCNET_PACKET_SAMPLES);
    lpcnet_decoder_destroy(st);
    return 0;
}

void ref_decode_packet(float *features /*[4][36]*/, float *vq_mem /*[18]*/, const unsigned char *buf)
{
    decode_packet((float (*)[NB_TOTAL_FEATURES])features, vq_mem, buf);
}

/* One call of the reference frame network on a fresh state after `warm` warm-up frames: dumps conditioning + lpc. */
int ref_frame_network(const unsigned char *blob, int len, const float *features, int stride, int nframes,
                      float *gru_a_cond /*[nframes][1152]*/, float *gru_b_cond /*[nframes][48]*/, float *lpc /*[nframes][16]*/)
{
    int i;
    LPCNetState *st = lpcnet_create();
    if (lpcnet_load_model(st, blob, len) != 0) { lpcnet_destroy(st); return -1; }
    for (i = 0; i < nframes; i++)
        run_frame_network(st, gru_a_cond + (size_t)i * 3 * GRU_A_STATE_SIZE, gru_b_cond + (size_t)i * 3 * GRU_B_STATE_SIZE,
                          lpc + (size_t)i * LPC_ORDER, features + (size_t)i * stride);
    lpcnet_destroy(st);
    return 0;
}

/* ---- CPU baseline timing: `nthreads` independent streams, each synthesising the same nframes. ---- */
typedef struct {
    const unsigned char *blob; int len; const float *features; int stride; int nframes; short *pcm; int decode;
    const unsigned char *packets;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    if (j->decode) ref_decode_stream(j->blob, j->len, j->packets, j->nframes, j->pcm);
    else ref_synth_stream(j->blob, j->len, j->features, j->stride, j->nframes, j->pcm);
    return NULL;
}

/* returns wall seconds; features: [nthreads][nframes][stride]; pcm: [nthreads][nframes*160] */
double ref_time_synthesis(const unsigned char *blob, int len, const float *features, int stride, int nframes,
                          int nthreads, short *pcm)
{
    struct timespec t0, t1;
    pthread_t *th = malloc(sizeof(*th) * nthreads);
    job_t *jobs = malloc(sizeof(*jobs) * nthreads);
    int i;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; i < nthreads; i++) {
        jobs[i].blob = blob; jobs[i].len = len; jobs[i].stride = stride; jobs[i].nframes = nframes; jobs[i].decode = 0;
        jobs[i].features = features + (size_t)i * nframes * stride;
        jobs[i].pcm = pcm + (size_t)i * nframes * LPCNET_FRAME_SIZE;
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
