"""Single-stream latency through the drop-in API (BASELINE config 1): 100 frames, one lpcnet_synthesize call per 10 ms frame,
and the same second of audio as ONE call of the batched API with one stream.  Prints ms per frame / x real time."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import helpers as H
import lpcnet_b200
from fixtures import make_feature_batch

T = 104
f = make_feature_batch(range(1), T)[0]
L = lpcnet_b200.lib()
L.lpcnet_b200_set_default_model(H.blob("int8"), len(H.blob("int8")), H.LPC_GAMMA)
net = lpcnet_b200.LPCNet()
for t in range(4): net.synthesize(f[t])                 # warm-up (also the silent frames)
t0 = time.time()
for t in range(4, T): net.synthesize(f[t])
dt = time.time() - t0
print("drop-in lpcnet_synthesize, 1 stream: %.3f ms per 10 ms frame (host call to PCM in host memory) -> %.2fx real time" % (1e3 * dt / (T - 4), 0.01 * (T - 4) / dt))
b = lpcnet_b200.Batch(1, H.blob("int8"), lpc_gamma=H.LPC_GAMMA)
b.synthesize(f[None, :4])
t0 = time.time()
b.synthesize(f[None, 4:])
dt = time.time() - t0
print("batched API, 1 stream, one call of %d frames: %.3f ms per frame -> %.2fx real time" % (T - 4, 1e3 * dt / (T - 4), 0.01 * (T - 4) / dt))
