/* TEST INFRASTRUCTURE.  The reference CLI (src/lpcnet_demo.c) also references the encoder and PLC entry points,
 * The encoder side is served by liblpcnet_b200.so itself; only the PLC (needs the PLC model) and two internal helpers the
 * demo's `-addlpc` mode calls are outside it.  These stubs let the UNTOUCHED demo source link against liblpcnet_b200.so so
 * that its -features, -encode, -synthesis and -decode modes exercise the drop-in API; the PLC / addlpc modes abort. */
#include <stdio.h>
#include <stdlib.h>
#include "lpcnet.h"
static void na(const char *f) { fprintf(stderr, "%s: not part of the B200 engine\n", f); exit(2); }
LPCNetPLCState *lpcnet_plc_create(int options) { (void)options; na("lpcnet_plc_create"); return 0; }
void lpcnet_plc_destroy(LPCNetPLCState *st) { (void)st; }
int lpcnet_plc_update(LPCNetPLCState *st, short *pcm) { (void)st; (void)pcm; na("lpcnet_plc_update"); return -1; }
int lpcnet_plc_conceal(LPCNetPLCState *st, short *pcm) { (void)st; (void)pcm; na("lpcnet_plc_conceal"); return -1; }
void lpcnet_plc_fec_add(LPCNetPLCState *st, const float *features) { (void)st; (void)features; na("lpcnet_plc_fec_add"); }
void lpcnet_plc_fec_clear(LPCNetPLCState *st) { (void)st; }
int lpcnet_plc_load_model(LPCNetPLCState *st, const unsigned char *data, int len) { (void)st; (void)data; (void)len; na("lpcnet_plc_load_model"); return -1; }
void lpc_weighting(float *lpc, float gamma) { (void)lpc; (void)gamma; na("lpc_weighting"); }
float lpc_from_cepstrum(float *lpc, const float *cepstrum) { (void)lpc; (void)cepstrum; na("lpc_from_cepstrum"); return 0; }
