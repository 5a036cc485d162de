/* lpcnet.h — public C API of the B200-native LPCNet synthesis engine.
 *
 * Drop-in for the synthesis / decoder / encoder / feature-extraction / model-loading part of xiph/LPCNet's public
 * header (reference: include/lpcnet.h).  Every prototype below has the same name, argument list, return type and
 * meaning as the reference declaration cited next to it, so `src/lpcnet_demo.c -features/-encode/-synthesis/-decode`
 * and any embedding application re-link against liblpcnet_b200.so unchanged.  The PLC entry points of the reference
 * header (include/lpcnet.h:191-212) are not exported: they need the PLC model (plc_data), which is outside the path; what
 * the PLC calls on the synthesis side (preload, tail, deferred frame network, state copies) is in lpcnet_b200.h.
 *
 * All state lives on the GPU.  The opaque structs below only hold a handle; `*_get_size()` is the size of that
 * handle.  There is no CPU fallback: if no CUDA device is usable `*_create` returns NULL, `*_init` returns -1
 * and the `synthesize/decode` calls emit zeros and latch an error readable with lpcnet_b200_last_error()
 * (include/lpcnet_b200.h).
 */
#ifndef _LPCNET_H_
#define _LPCNET_H_

#ifndef LPCNET_EXPORT
# if defined(__GNUC__)
#  define LPCNET_EXPORT __attribute__ ((visibility ("default")))
# else
#  define LPCNET_EXPORT
# endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define NB_FEATURES 20                 /* reference include/lpcnet.h:45 */
#define NB_TOTAL_FEATURES 36           /* :46 */
#define LPCNET_COMPRESSED_SIZE 8       /* :49  bytes per 40 ms packet */
#define LPCNET_PACKET_SAMPLES (4*160)  /* :50 */
#define LPCNET_FRAME_SIZE (160)        /* :51 */

typedef struct LPCNetState LPCNetState;
typedef struct LPCNetDecState LPCNetDecState;
typedef struct LPCNetEncState LPCNetEncState;

/* ---- encoder / feature extraction (PCM -> 1.6 kb/s packets, PCM -> feature frames) ---- */
LPCNET_EXPORT int lpcnet_encoder_get_size(void);                         /* ref :103 */
LPCNET_EXPORT int lpcnet_encoder_init(LPCNetEncState *st);               /* ref :112 returns 0 */
LPCNET_EXPORT LPCNetEncState *lpcnet_encoder_create(void);               /* ref :117 */
LPCNET_EXPORT void lpcnet_encoder_destroy(LPCNetEncState *st);           /* ref :122 */
/* pcm: LPCNET_PACKET_SAMPLES shorts in, buf: LPCNET_COMPRESSED_SIZE bytes out; returns 0 */
LPCNET_EXPORT int lpcnet_encode(LPCNetEncState *st, const short *pcm, unsigned char *buf);   /* ref :130 */
/* 640 samples -> four unquantised feature vectors */
LPCNET_EXPORT int lpcnet_compute_features(LPCNetEncState *st, const short *pcm, float features[4][NB_TOTAL_FEATURES]);   /* ref :138 */
/* LPCNET_FRAME_SIZE samples -> one feature vector (what `lpcnet_demo -features` writes per frame) */
LPCNET_EXPORT int lpcnet_compute_single_frame_features(LPCNetEncState *st, const short *pcm, float features[NB_TOTAL_FEATURES]);        /* ref :146 */
LPCNET_EXPORT int lpcnet_compute_single_frame_features_float(LPCNetEncState *st, const float *pcm, float features[NB_TOTAL_FEATURES]);  /* ref :155 */

/* ---- decoder (1.6 kb/s packets -> PCM) ---- */
LPCNET_EXPORT int lpcnet_decoder_get_size(void);                         /* ref :67  */
LPCNET_EXPORT int lpcnet_decoder_init(LPCNetDecState *st);               /* ref :76  returns 0 */
LPCNET_EXPORT LPCNetDecState *lpcnet_decoder_create(void);               /* ref :83  */
LPCNET_EXPORT void lpcnet_decoder_destroy(LPCNetDecState *st);           /* ref :88  */
/* buf: LPCNET_COMPRESSED_SIZE bytes in, pcm: LPCNET_PACKET_SAMPLES shorts out; returns 0 */
LPCNET_EXPORT int lpcnet_decode(LPCNetDecState *st, const unsigned char *buf, short *pcm);   /* ref :96 */

/* ---- synthesis (feature frames -> PCM) ---- */
LPCNET_EXPORT void lpcnet_reset(LPCNetState *lpcnet);                    /* ref :78  */
LPCNET_EXPORT int lpcnet_get_size(void);                                 /* ref :160 */
LPCNET_EXPORT int lpcnet_init(LPCNetState *st);                          /* ref :169 returns 0 */
LPCNET_EXPORT LPCNetState *lpcnet_create(void);                          /* ref :174 */
LPCNET_EXPORT void lpcnet_destroy(LPCNetState *st);                      /* ref :179 */
/* One call consumes ONE feature vector (first NB_FEATURES floats used) and produces N samples
 * (N = LPCNET_FRAME_SIZE in the demo/decoder).  The first two calls after a reset output silence. */
LPCNET_EXPORT void lpcnet_synthesize(LPCNetState *st, const float *features, short *output, int N);  /* ref :188 */

/* ---- model ingest: "DNNw" weight blob (reference src/write_lpcnet_weights.c) ---- */
/* Returns 0 on success, -1 if the blob does not describe a valid model (same contract as ref :214 / lpcnet.c:202). */
LPCNET_EXPORT int lpcnet_load_model(LPCNetState *st, const unsigned char *data, int len);            /* ref :214 */

#ifdef __cplusplus
}
#endif
#endif
