// lpcnet_api.cu — the reference's own public C API (include/lpcnet.h) as a thin host over the batched engine.
//
// Each LPCNetState / LPCNetDecState is a batch of ONE stream on device 0 (env LPCNET_B200_DEVICE overrides).
// Reference counterparts: src/lpcnet.c:169-224 (get_size/init/create/destroy/load_model/reset),
// :273-281 (lpcnet_synthesize), :285-319 (decoder wrappers).  Differences forced by the GPU:
//   * the opaque structs only hold a handle, device resources live in a registry that is drained at exit
//     (the reference API has `*_init` on caller memory with no matching deinit);
//   * the reference compiles its model and VQ codebooks in (nnet_data.c / ceps_codebooks.c); here they come from
//     lpcnet_load_model(), lpcnet_b200_set_default_model/_codebooks() or the LPCNET_B200_MODEL / _CODEBOOKS files;
//   * lpcnet_synthesize is void: on failure it emits zeros and latches lpcnet_b200_last_error().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "engine.h"
#include "../../include/lpcnet.h"
#include "../../include/lpcnet_b200.h"

using lpcnet_b200::set_error;

struct LPCNetState { uint32_t magic; uint32_t flags; LPCNetB200Batch *batch; };
struct LPCNetDecState { LPCNetState lpcnet_state; };   // same first-member layout as src/lpcnet_private.h:50-53
struct LPCNetEncState { uint32_t magic; uint32_t flags; LPCNetB200EncBatch *enc; };

namespace {
const uint32_t MAGIC = 0x4C50424Eu;   // "LPBN"
std::mutex g_mu;
std::vector<LPCNetB200Batch *> g_registry;
std::vector<LPCNetB200EncBatch *> g_enc_registry;
std::vector<unsigned char> g_model;
std::vector<float> g_codebooks;
float g_gamma = -1.0f;      // <= 0: not given -> the blob's metadata record, else 1 (lpcnet_b200_batch_create)
bool g_env_checked = false, g_atexit = false;

void drain()
{
    std::lock_guard<std::mutex> l(g_mu);
    for (auto *b : g_registry) lpcnet_b200_batch_destroy(b);
    for (auto *e : g_enc_registry) lpcnet_b200_enc_destroy(e);
    g_registry.clear(); g_enc_registry.clear();
}

bool read_file(const char *path, std::vector<unsigned char> &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? sz : 0);
    bool ok = sz > 0 && fread(out.data(), 1, sz, f) == (size_t)sz;
    fclose(f);
    return ok;
}
void check_env()     // caller holds g_mu
{
    if (g_env_checked) return;
    g_env_checked = true;
    if (const char *g = getenv("LPCNET_B200_LPC_GAMMA")) g_gamma = (float)atof(g);
    if (g_model.empty()) if (const char *p = getenv("LPCNET_B200_MODEL")) read_file(p, g_model);
    if (g_codebooks.empty()) if (const char *p = getenv("LPCNET_B200_CODEBOOKS")) {
        std::vector<unsigned char> raw;
        if (read_file(p, raw) && raw.size() % 4 == 0) { g_codebooks.resize(raw.size() / 4); memcpy(g_codebooks.data(), raw.data(), raw.size()); }
    }
}
int device_id() { const char *d = getenv("LPCNET_B200_DEVICE"); return d ? atoi(d) : 0; }

// Remove a batch from the registry; true if it was registered (i.e. it is alive and owned by the caller's state).  A state
// whose batch was already released — by drain() at exit, or because the caller's memory only LOOKS initialised — fails the
// look-up and must not free it again.
bool forget(LPCNetB200Batch *b)
{
    std::lock_guard<std::mutex> l(g_mu);
    for (size_t i = 0; i < g_registry.size(); i++) if (g_registry[i] == b) { g_registry.erase(g_registry.begin() + i); return true; }
    return false;
}
void release(LPCNetState *st)
{
    if (st->magic == MAGIC && st->batch && forget(st->batch)) lpcnet_b200_batch_destroy(st->batch);
    st->batch = nullptr;
}
int attach_model(LPCNetState *st, const unsigned char *blob, int len, float gamma)
{
    LPCNetB200Batch *nb = lpcnet_b200_batch_create(1, blob, len, gamma, device_id());
    if (!nb) return -1;
    {
        std::lock_guard<std::mutex> l(g_mu);
        if (!g_codebooks.empty()) lpcnet_b200_batch_set_codebooks(nb, g_codebooks.data(), g_codebooks.size());
        g_registry.push_back(nb);
        if (!g_atexit) { atexit(drain); g_atexit = true; }
    }
    release(st);
    st->batch = nb;
    return 0;
}
}  // namespace

extern "C" {

int lpcnet_b200_set_default_model(const unsigned char *blob, int len, float lpc_gamma)
{
    if (!blob || len <= 0) { set_error("default model: empty blob"); return -1; }
    std::lock_guard<std::mutex> l(g_mu);
    g_model.assign(blob, blob + len); g_gamma = lpc_gamma; g_env_checked = true;
    return 0;
}
int lpcnet_b200_set_default_codebooks(const float *cb, size_t n_floats)
{
    if (!cb || n_floats != 3 * 1024 * 17 + 4096 * 18) { set_error("default codebooks: wrong size"); return -1; }
    std::lock_guard<std::mutex> l(g_mu);
    g_codebooks.assign(cb, cb + n_floats);
    return 0;
}

int lpcnet_get_size(void) { return (int)sizeof(LPCNetState); }

int lpcnet_init(LPCNetState *st)
{
    if (!st) return -1;
    // re-initialising a live state (the reference allows lpcnet_init on the same memory again, src/lpcnet.c:184): release
    // the device batch it owns first.  Uninitialised caller memory that happens to carry the magic is harmless: its batch
    // pointer is not in the registry.
    release(st);
    st->magic = MAGIC; st->flags = 0; st->batch = nullptr;
    std::vector<unsigned char> model; float gamma;
    { std::lock_guard<std::mutex> l(g_mu); check_env(); model = g_model; gamma = g_gamma; }
    if (model.empty()) return 0;                 // like a USE_WEIGHTS_FILE build: model arrives via lpcnet_load_model (lpcnet.c:192-196)
    return attach_model(st, model.data(), (int)model.size(), gamma) == 0 ? 0 : -1;
}

LPCNetState *lpcnet_create(void)
{
    LPCNetState *st = (LPCNetState *)calloc(1, sizeof(LPCNetState));
    if (!st) return nullptr;
    if (lpcnet_b200_device_count() <= 0) { set_error("lpcnet_create: no CUDA device (no CPU fallback)"); free(st); return nullptr; }
    lpcnet_init(st);
    return st;
}

void lpcnet_destroy(LPCNetState *st)
{
    if (!st) return;
    release(st);
    free(st);
}

/* Explicit counterpart of lpcnet_init()/lpcnet_decoder_init() on caller-owned memory (the reference needs none because
 * its state owns no resources): frees the device batch now instead of at process exit. */
void lpcnet_b200_deinit(LPCNetState *st) { if (st) release(st); }

int lpcnet_load_model(LPCNetState *st, const unsigned char *data, int len)
{
    if (!st || st->magic != MAGIC) { set_error("lpcnet_load_model: state not initialised"); return -1; }
    float gamma;
    { std::lock_guard<std::mutex> l(g_mu); check_env(); gamma = g_gamma; }
    return attach_model(st, data, len, gamma) == 0 ? 0 : -1;
}

void lpcnet_reset(LPCNetState *st)
{
    if (st && st->magic == MAGIC && st->batch) lpcnet_b200_batch_reset(st->batch);
}

void lpcnet_synthesize(LPCNetState *st, const float *features, short *output, int N)
{
    if (N <= 0 || !output) return;
    if (!st || st->magic != MAGIC || !st->batch) {
        set_error("lpcnet_synthesize: no model loaded");
        memset(output, 0, sizeof(short) * N);
        return;
    }
    if (lpcnet_b200_batch_synthesize(st->batch, features, 1, NB_FEATURES, N, output) != 0) memset(output, 0, sizeof(short) * N);
}

int lpcnet_decoder_get_size(void) { return (int)sizeof(LPCNetDecState); }

int lpcnet_decoder_init(LPCNetDecState *st)
{
    if (!st) return -1;
    lpcnet_init(&st->lpcnet_state);              // (releases a batch a previous init of this memory created)
    return 0;
}

LPCNetDecState *lpcnet_decoder_create(void)
{
    LPCNetDecState *st = (LPCNetDecState *)calloc(1, sizeof(LPCNetDecState));
    if (!st) return nullptr;
    if (lpcnet_b200_device_count() <= 0) { set_error("lpcnet_decoder_create: no CUDA device (no CPU fallback)"); free(st); return nullptr; }
    lpcnet_decoder_init(st);
    return st;
}

void lpcnet_decoder_destroy(LPCNetDecState *st)
{
    if (!st) return;
    release(&st->lpcnet_state);
    free(st);
}

int lpcnet_decode(LPCNetDecState *st, const unsigned char *buf, short *pcm)
{
    if (!pcm) return -1;
    if (!st || st->lpcnet_state.magic != MAGIC || !st->lpcnet_state.batch) {
        set_error("lpcnet_decode: no model loaded");
        memset(pcm, 0, sizeof(short) * LPCNET_PACKET_SAMPLES);
        return -1;
    }
    if (lpcnet_b200_batch_decode(st->lpcnet_state.batch, buf, 1, pcm) != 0) { memset(pcm, 0, sizeof(short) * LPCNET_PACKET_SAMPLES); return -1; }
    return 0;
}


// ---- encoder / feature extraction: a batch of ONE analysis stream behind the reference's LPCNetEncState API (lpcnet_enc.c:466-486,882-933) ----
static void enc_release(LPCNetEncState *st)
{
    if (st->magic == MAGIC && st->enc) {
        bool mine = false;
        { std::lock_guard<std::mutex> l(g_mu); for (size_t i = 0; i < g_enc_registry.size(); i++) if (g_enc_registry[i] == st->enc) { g_enc_registry.erase(g_enc_registry.begin() + i); mine = true; break; } }
        if (mine) lpcnet_b200_enc_destroy(st->enc);
    }
    st->enc = nullptr;
}
int lpcnet_encoder_get_size(void) { return (int)sizeof(LPCNetEncState); }
int lpcnet_encoder_init(LPCNetEncState *st)
{
    if (!st) return -1;
    enc_release(st);
    st->magic = MAGIC; st->flags = 0; st->enc = nullptr;
    LPCNetB200EncBatch *e = lpcnet_b200_enc_create(1, device_id());
    if (!e) return -1;
    std::lock_guard<std::mutex> l(g_mu);
    check_env();
    if (!g_codebooks.empty()) lpcnet_b200_enc_set_codebooks(e, g_codebooks.data(), g_codebooks.size());
    g_enc_registry.push_back(e);
    if (!g_atexit) { atexit(drain); g_atexit = true; }
    st->enc = e;
    return 0;
}
LPCNetEncState *lpcnet_encoder_create(void)
{
    LPCNetEncState *st = (LPCNetEncState *)calloc(1, sizeof(LPCNetEncState));
    if (!st) return nullptr;
    if (lpcnet_encoder_init(st) != 0) { free(st); return nullptr; }
    return st;
}
void lpcnet_encoder_destroy(LPCNetEncState *st)
{
    if (!st) return;
    enc_release(st);
    free(st);
}
static LPCNetB200EncBatch *enc_of(LPCNetEncState *st, const char *who)
{
    if (!st || st->magic != MAGIC || !st->enc) { set_error("%s: encoder state not initialised", who); return nullptr; }
    return st->enc;
}
int lpcnet_encode(LPCNetEncState *st, const short *pcm, unsigned char *buf)
{
    LPCNetB200EncBatch *e = enc_of(st, "lpcnet_encode");
    return e && lpcnet_b200_enc_encode(e, pcm, 1, buf) == 0 ? 0 : -1;
}
int lpcnet_compute_features(LPCNetEncState *st, const short *pcm, float features[4][NB_TOTAL_FEATURES])
{
    LPCNetB200EncBatch *e = enc_of(st, "lpcnet_compute_features");
    return e && lpcnet_b200_enc_compute_features4(e, pcm, 1, &features[0][0]) == 0 ? 0 : -1;
}
int lpcnet_compute_single_frame_features(LPCNetEncState *st, const short *pcm, float features[NB_TOTAL_FEATURES])
{
    LPCNetB200EncBatch *e = enc_of(st, "lpcnet_compute_single_frame_features");
    return e && lpcnet_b200_enc_compute_features(e, pcm, 1, features) == 0 ? 0 : -1;
}
int lpcnet_compute_single_frame_features_float(LPCNetEncState *st, const float *pcm, float features[NB_TOTAL_FEATURES])
{
    LPCNetB200EncBatch *e = enc_of(st, "lpcnet_compute_single_frame_features_float");
    return e && lpcnet_b200_enc_compute_features_float(e, pcm, 1, features) == 0 ? 0 : -1;
}

}  // extern "C"
