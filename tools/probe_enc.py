"""Device-side timing probe of the analysis side (not a BASELINE metric): frames/s of batched feature extraction and packets/s of the
encoder for a few batch sizes, inputs/outputs resident on the device (CUDA events through the library's stream helpers are not exposed for
the encoder batch, so the host clock brackets a synchronous device-pointer call; the work per call is >= 10 ms)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import helpers as H
import lpcnet_b200
from fixtures import make_pcm_batch

L = lpcnet_b200.lib()
T = 100
base = make_pcm_batch(range(64), T)
for n in [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]:
    pcm = np.ascontiguousarray(np.tile(base, ((n + 63) // 64, 1))[:n])
    e = lpcnet_b200.EncBatch(n, codebooks=H.codebooks())
    d_pcm = L.lpcnet_b200_device_alloc(pcm.nbytes)
    d_f = L.lpcnet_b200_device_alloc(n * T * 36 * 4)
    d_p = L.lpcnet_b200_device_alloc(n * (T // 4) * 8)
    L.lpcnet_b200_memcpy_h2d(d_pcm, pcm.ctypes.data, pcm.nbytes)
    for it in range(2):
        t0 = time.time()
        assert L.lpcnet_b200_enc_compute_features_device(e._h, d_pcm, T, d_f, None) == 0
        tf = time.time() - t0
    e.reset()
    for it in range(2):
        t0 = time.time()
        assert L.lpcnet_b200_enc_encode_device(e._h, d_pcm, T // 4, d_p, None) == 0
        te = time.time() - t0
    print("n=%5d: features %.1f ms for %d frames -> %.3e frames/s (%.0fx real time per stream); encode %.1f ms for %d packets -> %.3e packets/s (%.0fx real time per stream)"
          % (n, tf * 1e3, n * T, n * T / tf, T * 0.01 / tf, te * 1e3, n * T // 4, n * (T // 4) / te, T * 0.01 / te), flush=True)
    for p in (d_pcm, d_f, d_p):
        L.lpcnet_b200_device_free(p)
    e.close()
