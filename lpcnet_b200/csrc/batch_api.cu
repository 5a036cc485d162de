// batch_api.cu — C-ABI of include/lpcnet_b200.h: batches of independent streams stepped in lockstep.
//
// Per call: [H2D features] -> (per chunk of <= CHUNK frames) frame network + LPC kernels -> persistent per-sample
// kernel -> [D2H PCM].  The per-stream state mirrors the resettable part of struct LPCNetState
// (reference src/lpcnet_private.h:28-48) in structure-of-arrays form (stream index fastest) so that lane==stream
// accesses coalesce.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "engine.h"
#include "devmath.cuh"
#include "../../include/lpcnet_b200.h"

namespace lpcnet_b200 { const char *get_error(); }
using namespace lpcnet_b200;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)

static const int CHUNK = 16;        // frames of conditioning buffered per sample-kernel launch

struct LPCNetB200Batch {
    int device, n;
    DeviceModel model;
    FrameState fs;
    // sample-rate state
    float *hA, *hB, *last_sig, *deemph; int *last_exc; uint32_t *rng;
    int frame_count;
    // work buffers
    float *condA, *condB, *lpc_raw;
    float *d_features; size_t d_features_cap;     // staging for host-pointer calls / decoder output
    short *d_pcm; size_t d_pcm_cap;
    uint8_t *d_packets; size_t d_packets_cap;
    cudaStream_t stream;
    // Ordering of the engine state across CUDA streams: every call that enqueues work records `order_ev` on the stream it
    // used; the next call makes ITS stream wait for that event first (and host-side readers synchronise on it).  A caller
    // may therefore pass a different cuda_stream to every `_device` call, or mix them with the host-pointer calls.
    cudaEvent_t order_ev; cudaStream_t last_stream; bool has_order;
    bool env_exact_cvt, env_float_lane_stream;    // LPCNET_B200_EXACT_CVT / _FLOAT_LANE_STREAM, read once at create time (tests)
    cudaEvent_t ev0, ev1;                         // user timer (lpcnet_b200_batch_timer_*)
    std::vector<cudaEvent_t> *kev;                // event pairs around every per-sample kernel launch of the last call
    int kev_used;
    int last_launches;
};

static int order_sync(LPCNetB200Batch *b);
// grow a staging buffer; work on any stream may still be using the old one, so wait for it before freeing
static int ensure(LPCNetB200Batch *b, void **p, size_t *cap, size_t bytes)
{
    if (*cap >= bytes) return 0;
    if (order_sync(b)) return -1;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes));
    *cap = bytes;
    return 0;
}

// stream hand-over (see LPCNetB200Batch::order_ev)
static int order_enter(LPCNetB200Batch *b, cudaStream_t st)
{
    if (b->has_order && b->last_stream != st) CK(cudaStreamWaitEvent(st, b->order_ev, 0));
    return 0;
}
static int order_leave(LPCNetB200Batch *b, cudaStream_t st)
{
    CK(cudaEventRecord(b->order_ev, st));
    b->last_stream = st; b->has_order = true;
    return 0;
}
static int order_sync(LPCNetB200Batch *b)       // host-side readers / writers of the state: wait for whatever ran last
{
    if (b->has_order) CK(cudaEventSynchronize(b->order_ev));
    return 0;
}

// The RCPPS look-up of the per-sample kernel forms its table address as `table | index` (devmath.cuh), which needs the
// table 8 KB-aligned in the SHARED WINDOW; the image offsets assume dynamic shared memory starts at SMEM_RESERVED.
// Probed once per device at create time so that a driver with a different reservation fails here with a message,
// not with a trap inside the kernel.
__global__ void smem_base_probe_kernel(uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t probe_smem[];
    *out = (uint32_t)__cvta_generic_to_shared(probe_smem);
}
static int probe_smem_base(uint32_t *base)
{
    uint32_t *d = nullptr;
    CK(cudaMalloc(&d, 4));
    CK(cudaFuncSetAttribute(smem_base_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_base_probe_kernel<<<1, 32, 227 * 1024>>>(d);
    cudaError_t e = cudaMemcpy(base, d, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("shared-memory probe failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

#ifdef LPCNET_TRACE
static long long *g_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int lpcnet_b200_debug_read_trace(long long *out)
{
    if (!g_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, g_trace, 8 * 32 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

int lpcnet_b200_version(void) { return 100; }
const char *lpcnet_b200_last_error(void) { return get_error(); }

int lpcnet_b200_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

// kiss99_srand(&rng, "LPCNet", 6): reference src/kiss99.c:32-57 called from lpcnet_reset (src/lpcnet.c:176-181).
static void kiss99_seed(uint32_t st[4])
{
    const unsigned char *d = (const unsigned char *)"LPCNet"; const int nd = 6;
    uint32_t z = 362436069u, w = 521288629u, jsr = 123456789u, jcong = 380116160u;
    int i;
    for (i = 3; i < nd; i += 4) {
        z ^= d[i - 3]; w ^= d[i - 2]; jsr ^= d[i - 1]; jcong ^= d[i];
        uint32_t znew = 36969u * (z & 0xFFFF) + (z >> 16), wnew = 18000u * (w & 0xFFFF) + (w >> 16);
        uint32_t shr3 = jsr ^ (jsr << 13); shr3 ^= shr3 >> 17; shr3 ^= shr3 << 5;
        z = znew; w = wnew; jsr = shr3; jcong = 69069u * jcong + 1234567u;
    }
    if (i - 3 < nd) z ^= d[i - 3];
    if (i - 2 < nd) w ^= d[i - 2];
    if (i - 1 < nd) jsr ^= d[i - 1];
    if (z == 0 || z == 0x9068FFFF) z++;
    if (w == 0 || w == 0x464FFFFF) w++;
    if (jsr == 0) jsr++;
    st[0] = z; st[1] = w; st[2] = jsr; st[3] = jcong;
}

int lpcnet_b200_batch_reset(LPCNetB200Batch *b)
{
    if (!b) { set_error("null batch"); return -1; }
    CK(cudaSetDevice(b->device));
    if (order_enter(b, b->stream)) return -1;
    const size_t n = b->n;
    CK(cudaMemsetAsync(b->hA, 0, sizeof(float) * NA * n, b->stream));
    CK(cudaMemsetAsync(b->hB, 0, sizeof(float) * NB * n, b->stream));
    CK(cudaMemsetAsync(b->last_sig, 0, sizeof(float) * LPC_ORDER * n, b->stream));
    CK(cudaMemsetAsync(b->deemph, 0, sizeof(float) * n, b->stream));
    CK(cudaMemsetAsync(b->fs.conv1_state, 0, sizeof(float) * 2 * FRAME_IN * n, b->stream));
    CK(cudaMemsetAsync(b->fs.conv2_state, 0, sizeof(float) * 2 * COND * n, b->stream));
    CK(cudaMemsetAsync(b->fs.lpc_carry, 0, sizeof(float) * 2 * LPC_ORDER * n, b->stream));
    CK(cudaMemsetAsync(b->fs.vq_mem, 0, sizeof(float) * NB_BANDS * n, b->stream));
    uint32_t seed[4];
    kiss99_seed(seed);
    std::vector<uint32_t> r(4 * n);
    for (int k = 0; k < 4; k++) for (size_t s = 0; s < n; s++) r[k * n + s] = seed[k];
    std::vector<int> le(n, 128);                       // last_exc = lin2ulaw(0.f) = 128 (lpcnet.c:180)
    CK(cudaMemcpyAsync(b->rng, r.data(), sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(b->last_exc, le.data(), sizeof(int) * n, cudaMemcpyHostToDevice, b->stream));
    if (order_leave(b, b->stream)) return -1;
    CK(cudaStreamSynchronize(b->stream));
    b->frame_count = 0;
    return 0;
}

LPCNetB200Batch *lpcnet_b200_batch_create(int n_streams, const unsigned char *blob, int blob_len, float lpc_gamma, int device)
{
    if (n_streams <= 0) { set_error("n_streams must be positive"); return nullptr; }
    int cnt = lpcnet_b200_device_count();
    if (cnt <= 0) { set_error("no CUDA device available (this engine has no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= cnt) { set_error("device %d out of range (%d devices)", device, cnt); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    LPCNetB200Batch *b = (LPCNetB200Batch *)calloc(1, sizeof(*b));
    b->device = device; b->n = n_streams;
    b->kev = new std::vector<cudaEvent_t>();
    if (model_load(&b->model, blob, blob_len, lpc_gamma) != 0) { delete b->kev; free(b); return nullptr; }
    const size_t n = n_streams;
    bool ok = true;
    auto al = [&](void **p, size_t bytes) { if (ok && cudaMalloc(p, bytes) != cudaSuccess) ok = false; };
    al((void **)&b->hA, sizeof(float) * NA * n); al((void **)&b->hB, sizeof(float) * NB * n);
    al((void **)&b->last_sig, sizeof(float) * LPC_ORDER * n); al((void **)&b->deemph, sizeof(float) * n);
    al((void **)&b->last_exc, sizeof(int) * n); al((void **)&b->rng, sizeof(uint32_t) * 4 * n);
    al((void **)&b->fs.conv1_state, sizeof(float) * 2 * FRAME_IN * n); al((void **)&b->fs.conv2_state, sizeof(float) * 2 * COND * n);
    al((void **)&b->fs.lpc_carry, sizeof(float) * 2 * LPC_ORDER * n); al((void **)&b->fs.vq_mem, sizeof(float) * NB_BANDS * n);
    al((void **)&b->condA, sizeof(float) * (size_t)CHUNK * n * 3 * NA); al((void **)&b->condB, sizeof(float) * (size_t)CHUNK * n * 3 * NB);
    al((void **)&b->lpc_raw, sizeof(float) * (size_t)(CHUNK + 2) * n * LPC_ORDER);
    if (ok && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    if (ok && (cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess)) ok = false;
    if (ok && cudaEventCreateWithFlags(&b->order_ev, cudaEventDisableTiming) != cudaSuccess) ok = false;
    b->env_exact_cvt = getenv("LPCNET_B200_EXACT_CVT") != nullptr;
    b->env_float_lane_stream = getenv("LPCNET_B200_FLOAT_LANE_STREAM") != nullptr;
    if (ok && !b->model.is_float) {
        uint32_t base = 0;
        if (probe_smem_base(&base)) { lpcnet_b200_batch_destroy(b); return nullptr; }
        if ((base + SM_IMAGE + IM_RCP) % 8192u != 0) {
            set_error("dynamic shared memory starts at window offset %u on this driver, the image layout assumes %u (engine.h SMEM_RESERVED)", base, SMEM_RESERVED);
            lpcnet_b200_batch_destroy(b); return nullptr;
        }
    }
    if (!ok) { set_error("device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); lpcnet_b200_batch_destroy(b); return nullptr; }
    if (lpcnet_b200_batch_reset(b) != 0) { lpcnet_b200_batch_destroy(b); return nullptr; }
    return b;
}

void lpcnet_b200_batch_destroy(LPCNetB200Batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    void *ptrs[] = {b->hA, b->hB, b->last_sig, b->deemph, b->last_exc, b->rng, b->fs.conv1_state, b->fs.conv2_state, b->fs.lpc_carry,
                    b->fs.vq_mem, b->condA, b->condB, b->lpc_raw, b->d_features, b->d_pcm, b->d_packets};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (b->kev) { for (cudaEvent_t e : *b->kev) cudaEventDestroy(e); delete b->kev; }
    model_free(&b->model);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->order_ev) { if (b->has_order) cudaEventSynchronize(b->order_ev); cudaEventDestroy(b->order_ev); }
    if (b->stream) cudaStreamDestroy(b->stream);
    free(b);
}

int lpcnet_b200_batch_streams(const LPCNetB200Batch *b) { return b ? b->n : 0; }
int lpcnet_b200_batch_is_float(const LPCNetB200Batch *b) { return b ? b->model.is_float : -1; }

int lpcnet_b200_batch_set_codebooks(LPCNetB200Batch *b, const float *cb, size_t n_floats)
{
    if (!b) { set_error("null batch"); return -1; }
    const size_t want = 3 * 1024 * 17 + 4096 * 18;
    if (!cb || n_floats != want) { set_error("codebooks: expected %zu floats", want); return -1; }
    CK(cudaSetDevice(b->device));
    if (!b->model.codebooks) CK(cudaMalloc((void **)&b->model.codebooks, want * sizeof(float)));
    CK(cudaMemcpy(b->model.codebooks, cb, want * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// Core: features and pcm are device pointers. `st` is the stream all work is enqueued on.
static int synth_device(LPCNetB200Batch *b, const float *d_feat, long long stream_stride, int frame_stride, int nframes,
                        int spf, short *d_pcm, cudaStream_t st, bool time_it)
{
    if (!b) { set_error("null batch"); return -1; }
    if (nframes <= 0) return 0;
    if (spf < 1 || spf > 65536) { set_error("samples_per_frame must be in 1..65536"); return -1; }
    if (frame_stride < NB_FEAT) { set_error("feature_stride must be >= 20"); return -1; }
    const int n = b->n;
    const long long pcm_stride = (long long)nframes * spf;
    int launches = 0;
    b->kev_used = 0;
    if (order_enter(b, st)) return -1;
    for (int c0 = 0; c0 < nframes; c0 += CHUNK) {
        const int nf = nframes - c0 < CHUNK ? nframes - c0 : CHUNK;
        launch_frame_network(b->model, b->fs, d_feat + (size_t)c0 * frame_stride, stream_stride, frame_stride, n, nf,
                             b->frame_count, b->condA, b->condB, b->lpc_raw, st);
        launches += 4;
        // frames whose post-increment frame_count is <= FEATURES_DELAY are silent and do not advance the sample
        // state (lpcnet.c:239-243)
        int silent = FEATURES_DELAY - b->frame_count;
        if (silent < 0) silent = 0;
        if (silent > nf) silent = nf;
        if (silent > 0)
            CK(cudaMemset2DAsync(d_pcm + (size_t)c0 * spf, pcm_stride * sizeof(short), 0, (size_t)silent * spf * sizeof(short), n, st));
        if (nf > silent) {
            SampleParams p = {};
            p.L = b->model.L; p.image = b->model.image;
            p.emb_sig = b->model.emb_sig; p.emb_pred = b->model.emb_pred; p.emb_exc = b->model.emb_exc; p.fcw = b->model.fcw;
            p.spc = streams_per_cta_for(n);
            p.fast_cvt = b->model.fast_cvt && !b->env_exact_cvt;   // env LPCNET_B200_EXACT_CVT: force the conversion-unit path (tests)
#ifdef LPCNET_TRACE
            { static long long *d_trace = nullptr; if (!d_trace) { CK(cudaMalloc(&d_trace, 8 * 32 * 8)); CK(cudaMemset(d_trace, 0, 8 * 32 * 8)); } p.trace = d_trace; g_trace = d_trace; }
#endif
            p.condA = b->condA + (size_t)silent * n * 3 * NA;
            p.condB = b->condB + (size_t)silent * n * 3 * NB;
            p.lpc_raw = b->lpc_raw + (size_t)silent * n * LPC_ORDER;
            p.gamma_pow = b->model.gamma_pow;
            p.hA = b->hA; p.hB = b->hB; p.last_sig = b->last_sig; p.deemph = b->deemph; p.last_exc = b->last_exc; p.rng = b->rng;
            p.pcm = d_pcm + (size_t)(c0 + silent) * spf;
            p.pcm_stream_stride = pcm_stride;
            p.n_streams = n; p.nframes = nf - silent; p.spf = spf;
            if (time_it) {      // event pair around the launch, resolved lazily by lpcnet_b200_batch_last_sample_kernel_ms (no sync here)
                while ((int)b->kev->size() < b->kev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); b->kev->push_back(e); }
                CK(cudaEventRecord((*b->kev)[b->kev_used], st));
            }
            if (b->model.is_float && p.spc <= FN_S && !b->env_float_lane_stream) {
                p.L = b->model.Ln; p.image = b->model.image_n;       // small batch: neuron-per-lane float kernel
                CK(launch_sample_kernel_f32n(p, st));
            } else
                CK(b->model.is_float ? launch_sample_kernel_f32(p, st) : launch_sample_kernel(p, st));
            launches += 1;
            if (time_it) { CK(cudaEventRecord((*b->kev)[b->kev_used + 1], st)); b->kev_used += 2; }
        }
        b->frame_count += nf;
        if (b->frame_count > 1000) b->frame_count = 1000;
    }
    CK(cudaGetLastError());
    if (order_leave(b, st)) return -1;
    b->last_launches = launches;
    return 0;
}

int lpcnet_b200_batch_synthesize_device(LPCNetB200Batch *b, const float *d_features, int nframes, int feature_stride,
                                        int samples_per_frame, short *d_pcm, void *cuda_stream)
{
    if (!b) { set_error("null batch"); return -1; }
    CK(cudaSetDevice(b->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    int r = synth_device(b, d_features, (long long)nframes * feature_stride, feature_stride, nframes, samples_per_frame, d_pcm, st, true);
    if (r == 0 && !cuda_stream) CK(cudaStreamSynchronize(st));
    return r;
}

int lpcnet_b200_batch_synthesize(LPCNetB200Batch *b, const float *features, int nframes, int feature_stride,
                                 int samples_per_frame, short *pcm)
{
    if (!b) { set_error("null batch"); return -1; }
    if (nframes <= 0) return 0;
    if (!features || !pcm) { set_error("null buffer"); return -1; }
    CK(cudaSetDevice(b->device));
    const size_t fbytes = sizeof(float) * (size_t)b->n * nframes * feature_stride;
    const size_t pbytes = sizeof(short) * (size_t)b->n * nframes * samples_per_frame;
    if (ensure(b, (void **)&b->d_features, &b->d_features_cap, fbytes)) return -1;
    if (ensure(b, (void **)&b->d_pcm, &b->d_pcm_cap, pbytes)) return -1;
    if (order_enter(b, b->stream)) return -1;       // the staging buffers may still be read by work on another stream
    CK(cudaMemcpyAsync(b->d_features, features, fbytes, cudaMemcpyHostToDevice, b->stream));
    if (order_leave(b, b->stream)) return -1;
    if (synth_device(b, b->d_features, (long long)nframes * feature_stride, feature_stride, nframes, samples_per_frame, b->d_pcm, b->stream, true)) return -1;
    CK(cudaMemcpyAsync(pcm, b->d_pcm, pbytes, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

static int decode_device(LPCNetB200Batch *b, const uint8_t *d_packets, int npackets, short *d_pcm, cudaStream_t st)
{
    if (!b->model.codebooks) { set_error("decode: no VQ codebooks loaded (lpcnet_b200_batch_set_codebooks)"); return -1; }
    const size_t fbytes = sizeof(float) * (size_t)b->n * npackets * 4 * NB_FEAT;
    if (ensure(b, (void **)&b->d_features, &b->d_features_cap, fbytes)) return -1;
    if (order_enter(b, st)) return -1;
    launch_decode_packets(b->model, b->fs, d_packets, b->n, npackets, b->d_features, st);
    if (order_leave(b, st)) return -1;
    int r = synth_device(b, b->d_features, (long long)npackets * 4 * NB_FEAT, NB_FEAT, npackets * 4, FRAME_SIZE, d_pcm, st, true);
    b->last_launches += 1;
    return r;
}

int lpcnet_b200_batch_decode_device(LPCNetB200Batch *b, const unsigned char *d_packets, int npackets, short *d_pcm, void *cuda_stream)
{
    if (!b) { set_error("null batch"); return -1; }
    if (npackets <= 0) return 0;
    CK(cudaSetDevice(b->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : b->stream;
    int r = decode_device(b, d_packets, npackets, d_pcm, st);
    if (r == 0 && !cuda_stream) CK(cudaStreamSynchronize(st));
    return r;
}

int lpcnet_b200_batch_decode(LPCNetB200Batch *b, const unsigned char *packets, int npackets, short *pcm)
{
    if (!b) { set_error("null batch"); return -1; }
    if (npackets <= 0) return 0;
    if (!packets || !pcm) { set_error("null buffer"); return -1; }
    CK(cudaSetDevice(b->device));
    const size_t kbytes = (size_t)b->n * npackets * 8, pbytes = sizeof(short) * (size_t)b->n * npackets * 640;
    if (ensure(b, (void **)&b->d_packets, &b->d_packets_cap, kbytes)) return -1;
    if (ensure(b, (void **)&b->d_pcm, &b->d_pcm_cap, pbytes)) return -1;
    if (order_enter(b, b->stream)) return -1;
    CK(cudaMemcpyAsync(b->d_packets, packets, kbytes, cudaMemcpyHostToDevice, b->stream));
    if (order_leave(b, b->stream)) return -1;
    if (decode_device(b, b->d_packets, npackets, b->d_pcm, b->stream)) return -1;
    CK(cudaMemcpyAsync(pcm, b->d_pcm, pbytes, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

float lpcnet_b200_batch_last_sample_kernel_ms(const LPCNetB200Batch *b, int *launches)
{
    if (!b) return 0.f;
    if (launches) *launches = b->last_launches;
    float total = 0.f;
    for (int i = 0; i + 1 < b->kev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventSynchronize((*b->kev)[i + 1]) != cudaSuccess) return -1.f;
        if (cudaEventElapsedTime(&ms, (*b->kev)[i], (*b->kev)[i + 1]) != cudaSuccess) return -1.f;
        total += ms;
    }
    return total;
}

// CUDA-event stopwatch on the engine's own stream (the stream every engine kernel is launched on when the caller
// passes cuda_stream == NULL): start records an event, stop records another, synchronises and returns the ms between.
int lpcnet_b200_batch_timer_start(LPCNetB200Batch *b)
{
    if (!b) { set_error("null batch"); return -1; }
    CK(cudaSetDevice(b->device));
    CK(cudaEventRecord(b->ev0, b->stream));
    return 0;
}
float lpcnet_b200_batch_timer_stop(LPCNetB200Batch *b)
{
    if (!b) return -1.f;
    float ms = 0.f;
    if (cudaEventRecord(b->ev1, b->stream) != cudaSuccess || cudaEventSynchronize(b->ev1) != cudaSuccess ||
        cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) { set_error("timer: %s", cudaGetErrorString(cudaGetLastError())); return -1.f; }
    return ms;
}
// Evict L2 between timed iterations: overwrite a scratch buffer larger than the 126 MB L2 on the engine's stream.
int lpcnet_b200_batch_flush_l2(LPCNetB200Batch *b)
{
    if (!b) { set_error("null batch"); return -1; }
    static void *scratch = nullptr; static int scratch_dev = -1;
    const size_t bytes = 256u << 20;
    CK(cudaSetDevice(b->device));
    if (!scratch || scratch_dev != b->device) { CK(cudaMalloc(&scratch, bytes)); scratch_dev = b->device; }
    CK(cudaMemsetAsync(scratch, 0x5a, bytes, b->stream));
    return 0;
}
int lpcnet_b200_batch_sync(LPCNetB200Batch *b)
{
    if (!b) { set_error("null batch"); return -1; }
    CK(cudaSetDevice(b->device));
    if (order_sync(b)) return -1;
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}
// Device memory helpers so that a plain-C / ctypes caller can keep inputs resident in HBM (no torch needed).
void *lpcnet_b200_device_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void lpcnet_b200_device_free(void *p) { if (p) cudaFree(p); }
int lpcnet_b200_memcpy_h2d(void *dst, const void *src, size_t bytes) { CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return 0; }
int lpcnet_b200_memcpy_d2h(void *dst, const void *src, size_t bytes) { CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return 0; }

int lpcnet_b200_batch_algorithmic_bytes(const LPCNetB200Batch *b, long *total, long *sparse_gemv)
{
    if (!b) { set_error("null batch"); return -1; }
    if (total) *total = b->model.algo_bytes_total;
    if (sparse_gemv) *sparse_gemv = b->model.algo_bytes_sparse;
    return 0;
}

int lpcnet_b200_batch_get_state(LPCNetB200Batch *b, int s, float *gru_a, float *gru_b, float *last_sig, int *misc, uint32_t *rng)
{
    if (!b || s < 0 || s >= b->n) { set_error("bad stream index"); return -1; }
    CK(cudaSetDevice(b->device));
    if (order_sync(b)) return -1;
    const size_t n = b->n;
    if (gru_a) CK(cudaMemcpy2D(gru_a, sizeof(float), b->hA + s, n * sizeof(float), sizeof(float), NA, cudaMemcpyDeviceToHost));
    if (gru_b) CK(cudaMemcpy2D(gru_b, sizeof(float), b->hB + s, n * sizeof(float), sizeof(float), NB, cudaMemcpyDeviceToHost));
    if (last_sig) CK(cudaMemcpy2D(last_sig, sizeof(float), b->last_sig + s, n * sizeof(float), sizeof(float), LPC_ORDER, cudaMemcpyDeviceToHost));
    if (misc) { CK(cudaMemcpy(&misc[0], b->last_exc + s, sizeof(int), cudaMemcpyDeviceToHost)); misc[1] = b->frame_count; }
    if (rng) CK(cudaMemcpy2D(rng, sizeof(uint32_t), b->rng + s, n * sizeof(uint32_t), sizeof(uint32_t), 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Test hook: run ONLY the frame-rate kernels on host features for a fresh batch state and return all taps.
// ga [n][nframes][1152], gb [n][nframes][48], lpc [n][nframes][16] (gamma-weighted, i.e. what the sample loop uses)
int lpcnet_b200_debug_frame_network(LPCNetB200Batch *b, const float *features, int nframes, int feature_stride,
                                                  float *ga, float *gb, float *lpc)
{
    if (!b) { set_error("null batch"); return -1; }
    if (nframes > CHUNK) { set_error("debug_frame_network: at most %d frames", CHUNK); return -1; }
    CK(cudaSetDevice(b->device));
    const int n = b->n;
    const size_t fbytes = sizeof(float) * (size_t)n * nframes * feature_stride;
    if (order_sync(b)) return -1;
    if (ensure(b, (void **)&b->d_features, &b->d_features_cap, fbytes)) return -1;
    CK(cudaMemcpyAsync(b->d_features, features, fbytes, cudaMemcpyHostToDevice, b->stream));
    launch_frame_network(b->model, b->fs, b->d_features, (long long)nframes * feature_stride, feature_stride, n, nframes,
                         b->frame_count, b->condA, b->condB, b->lpc_raw, b->stream);
    b->frame_count += nframes;
    std::vector<float> hA((size_t)nframes * n * 3 * NA), hB((size_t)nframes * n * 3 * NB), hl((size_t)(nframes + 2) * n * LPC_ORDER), gp(LPC_ORDER);
    CK(cudaMemcpyAsync(hA.data(), b->condA, hA.size() * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(hB.data(), b->condB, hB.size() * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(hl.data(), b->lpc_raw, hl.size() * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(gp.data(), b->model.gamma_pow, LPC_ORDER * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int s = 0; s < n; s++) for (int f = 0; f < nframes; f++) {
        memcpy(ga + ((size_t)s * nframes + f) * 3 * NA, &hA[((size_t)f * n + s) * 3 * NA], sizeof(float) * 3 * NA);
        memcpy(gb + ((size_t)s * nframes + f) * 3 * NB, &hB[((size_t)f * n + s) * 3 * NB], sizeof(float) * 3 * NB);
        for (int i = 0; i < LPC_ORDER; i++) lpc[((size_t)s * nframes + f) * LPC_ORDER + i] = hl[((size_t)f * n + s) * LPC_ORDER + i] * gp[i];
    }
    return 0;
}

// Test hook (host only, no CUDA): the shared-memory image and its run-time layout words
// layout[8] = {wA, metaA, wB, metaB, image_bytes, total_bytes, nblkA_padded, nblkB_padded}; also returns SM_IMAGE in layout[8].
int lpcnet_b200_debug_image(const unsigned char *blob, int len, unsigned char *out, size_t cap, uint32_t *layout)
{
    SmemLayout L;
    int r = debug_build_image(blob, len, out, cap, &L);
    if (r < 0) return r;
    layout[0] = L.wA; layout[1] = L.metaA; layout[2] = L.wB; layout[3] = L.metaB; layout[4] = L.image_bytes; layout[5] = L.total_bytes;
    layout[6] = L.nblkA_padded; layout[7] = L.nblkB_padded; layout[8] = L.sm_image;
    layout[9] = IM_PARA; layout[10] = IM_DIRA; layout[11] = IM_GRPA; layout[12] = IM_DIRB; layout[13] = IM_WBREC; layout[14] = IM_PARB;
    layout[15] = IM_FCW; layout[16] = NWC; layout[17] = GPW; layout[18] = *reinterpret_cast<const uint32_t *>(out + IM_FCWN); layout[19] = KPARTS;
    return r;
}

// Test hook: the two device implementations of the reference's _mm256_rcp_ps (table in memory / table-free arithmetic)
// on n host floats.
__global__ void debug_rcp_kernel(const float *x, float *out_table, float *out_arith, int n, const uint16_t *tab16)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_table[i] = rcp_emul(x[i], tab16);
    out_arith[i] = rcp_emul(x[i], RcpArith());
}
int lpcnet_b200_debug_rcp(LPCNetB200Batch *b, const float *x, float *out_table, float *out_arith, int n)
{
    if (!b || n <= 0) { set_error("debug_rcp: bad arguments"); return -1; }
    CK(cudaSetDevice(b->device));
    float *d = nullptr;
    CK(cudaMalloc(&d, (size_t)3 * n * sizeof(float)));
    CK(cudaMemcpy(d, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    debug_rcp_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(d, d + n, d + 2 * n, n, b->model.rcp16);
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaMemcpy(out_table, d + n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_arith, d + 2 * n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// Test hook (host only): image of the neuron-per-lane float kernel.  layout = {wA, metaA, wB, metaB, image_bytes, total_bytes,
// nblkA, nblkB, FN_IMAGE, FNI_NEUR, FNI_DIRA, FNI_PARA, FNI_DIRB, FNI_PARB, FNI_WBREC, FNI_FCW, dense flag}
int lpcnet_b200_debug_image_n(const unsigned char *blob, int len, unsigned char *out, size_t cap, uint32_t *layout)
{
    SmemLayout L;
    int r = debug_build_image_n(blob, len, out, cap, &L);
    if (r < 0) return r;
    layout[0] = L.wA; layout[1] = L.metaA; layout[2] = L.wB; layout[3] = L.metaB; layout[4] = L.image_bytes; layout[5] = L.total_bytes;
    layout[6] = L.nblkA_padded; layout[7] = L.nblkB_padded; layout[8] = FN_IMAGE; layout[9] = FNI_NEUR; layout[10] = FNI_DIRA; layout[11] = FNI_PARA;
    layout[12] = FNI_DIRB; layout[13] = FNI_PARB; layout[14] = FNI_WBREC; layout[15] = FNI_FCW; layout[16] = L.wBrecF;
    return r;
}

// Pinned host memory helpers for callers that want true async H2D/D2H (the benchmark's e2e leg).
void *lpcnet_b200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void lpcnet_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
