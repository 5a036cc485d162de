// blob_io.cu — the on-disk model format as first-class I/O: "DNNw" blob writer / lister and file helpers.
// Host code only (no kernels); lives in a .cu for build uniformity.
//
// Replaces (reference file:line):
//   write_weights                src/write_lpcnet_weights.c:47-67   (64-byte WeightHead + payload padded to 64 bytes)
//   parse_record / parse_weights src/parse_lpcnet_weights.c:37-76   (enumeration; the same acceptance rules)
// The reference keeps LPC_GAMMA / FEATURES_DELAY / END2END in the generated nnet_data.h (training_tf2/dump_lpcnet.py:306-329),
// i.e. outside the blob.  The writer can append them as one extra record, `lpcnet_b200_config` (float [4] = gamma, delay,
// end2end, format version); the reference's loader looks arrays up by name and ignores the rest, so such a blob still loads there.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "engine.h"
#include "../../include/lpcnet_b200.h"

using namespace lpcnet_b200;

namespace {
struct Head { char head[4]; int version, type, size, block_size; char name[44]; };     // WeightHead, src/nnet.h:54-61
static_assert(sizeof(Head) == 64, "WeightHead is one 64-byte block");
const char *kConfigName = "lpcnet_b200_config";

size_t record_bytes(int size) { return 64 + ((size_t)size + 63) / 64 * 64; }

void put_record(unsigned char *dst, const char *name, int type, int size, const void *data)
{
    Head h;
    memset(&h, 0, sizeof(h));
    memcpy(h.head, "DNNw", 4);
    h.version = 0; h.type = type; h.size = size; h.block_size = (size + 63) / 64 * 64;
    strncpy(h.name, name, sizeof(h.name) - 1);
    memcpy(dst, &h, 64);
    memcpy(dst + 64, data, (size_t)size);
    memset(dst + 64 + size, 0, (size_t)h.block_size - size);
}
}  // namespace

extern "C" {

long long lpcnet_b200_write_blob(const LPCNetB200Array *arrays, int count, const LPCNetB200Config *cfg, unsigned char *out, size_t cap)
{
    if (count < 0 || (count > 0 && !arrays)) { set_error("write_blob: bad arguments"); return -1; }
    size_t need = 0;
    for (int i = 0; i < count; i++) {
        const LPCNetB200Array &a = arrays[i];
        if (!a.name || !a.data || a.size <= 0) { set_error("write_blob: array %d is empty or unnamed", i); return -1; }
        if (strlen(a.name) > 43) { set_error("write_blob: name '%s' longer than 43 characters", a.name); return -1; }
        if (a.type < 0 || a.type > 2) { set_error("write_blob: array '%s' has unknown type %d", a.name, a.type); return -1; }
        if (cfg && !strcmp(a.name, kConfigName)) continue;      // replaced by the record built from cfg
        need += record_bytes(a.size);
    }
    if (cfg) {
        if (cfg->features_delay < 0 || cfg->features_delay > MAX_FEATURES_DELAY || !(cfg->lpc_gamma > 0.f) || cfg->end2end < 0) {
            set_error("write_blob: config record needs lpc_gamma > 0, features_delay 0..%d, end2end 0/1", MAX_FEATURES_DELAY); return -1;
        }
        need += record_bytes(16);
    }
    if (!out) return (long long)need;                          // size query
    if (cap < need) { set_error("write_blob: buffer of %zu bytes, %zu needed", cap, need); return -1; }
    unsigned char *p = out;
    for (int i = 0; i < count; i++) {
        const LPCNetB200Array &a = arrays[i];
        if (cfg && !strcmp(a.name, kConfigName)) continue;
        put_record(p, a.name, a.type, a.size, a.data);
        p += record_bytes(a.size);
    }
    if (cfg) {
        const float v[4] = {cfg->lpc_gamma, (float)cfg->features_delay, cfg->end2end ? 1.f : 0.f, 1.f};
        put_record(p, kConfigName, 0, 16, v);
        p += record_bytes(16);
    }
    return (long long)(p - out);
}

int lpcnet_b200_parse_blob(const unsigned char *blob, int len, LPCNetB200Array *arrays, int cap)
{
    if (!blob || len <= 0) { set_error("parse_blob: empty blob"); return -1; }
    int n = 0;
    const unsigned char *d = blob;
    while (len > 0) {
        if (len < 64) { set_error("parse_blob: truncated record header"); return -1; }
        const Head *h = reinterpret_cast<const Head *>(d);
        if (memcmp(h->head, "DNNw", 4) || h->version != 0) { set_error("parse_blob: record %d is not a DNNw v0 record", n); return -1; }
        if (h->size <= 0 || h->block_size < h->size || h->block_size > len - 64 || h->name[43] != 0) { set_error("parse_blob: record %d malformed", n); return -1; }
        if (arrays && n < cap) { arrays[n].name = h->name; arrays[n].type = h->type; arrays[n].size = h->size; arrays[n].data = d + 64; }
        n++;
        d += 64 + h->block_size; len -= 64 + h->block_size;
    }
    return n;
}

int lpcnet_b200_blob_config(const unsigned char *blob, int len, LPCNetB200Config *cfg)
{
    if (!cfg) { set_error("blob_config: null output"); return -1; }
    std::vector<LPCNetB200Array> a(4096);
    const int n = lpcnet_b200_parse_blob(blob, len, a.data(), (int)a.size());
    if (n < 0) return -1;
    for (int i = 0; i < n && i < (int)a.size(); i++)
        if (!strcmp(a[i].name, kConfigName) && a[i].size >= 12) {
            const float *v = reinterpret_cast<const float *>(a[i].data);
            cfg->lpc_gamma = v[0]; cfg->features_delay = (int)v[1]; cfg->end2end = v[2] != 0.f;
            return 1;
        }
    return 0;
}

int lpcnet_b200_write_blob_file(const char *path, const LPCNetB200Array *arrays, int count, const LPCNetB200Config *cfg)
{
    const long long need = lpcnet_b200_write_blob(arrays, count, cfg, nullptr, 0);
    if (need < 0) return -1;
    std::vector<unsigned char> buf((size_t)need);
    if (lpcnet_b200_write_blob(arrays, count, cfg, buf.data(), buf.size()) != need) return -1;
    FILE *f = path ? fopen(path, "wb") : nullptr;
    if (!f) { set_error("write_blob_file: cannot open '%s'", path ? path : "(null)"); return -1; }
    const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    if (fclose(f) != 0 || !ok) { set_error("write_blob_file: short write to '%s'", path); return -1; }
    return 0;
}

long long lpcnet_b200_read_file(const char *path, unsigned char *out, size_t cap)
{
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) { set_error("read_file: cannot open '%s'", path ? path : "(null)"); return -1; }
    fseek(f, 0, SEEK_END);
    const long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (out) {
        if ((long long)cap < sz || fread(out, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); set_error("read_file: buffer too small or short read"); return -1; }
    }
    fclose(f);
    return sz;
}

}  // extern "C"
