"""Timeline of CTA 0 (tuning build with -DLPCNET_TRACE): prints clock64 stamps of samples 200..207 relative to sample 200's start."""
import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import helpers as H
import lpcnet_b200
from lpcnet_b200 import api
from fixtures import make_feature_batch
n = 4096
base = make_feature_batch(range(64), 8)
f = base[np.arange(n) % 64]
b = lpcnet_b200.Batch(n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA)
b.synthesize(f[:, :4]); b.synthesize(f[:, 4:])
L = api.lib()
out = (ctypes.c_longlong * (256 + 8 * 32 * 16))()
L.lpcnet_b200_debug_read_trace.argtypes = [ctypes.c_void_p]
assert L.lpcnet_b200_debug_read_trace(out) == 0
allt = np.array(out, dtype=np.int64)
t = allt[:256].reshape(8, 32)
tw = allt[256:].reshape(8, 32, 16)
t0 = t[0, 0] if t[0, 0] else tw[0, 0, 0]
names = {0: "c:start", 1: "c:gemvA done", 2: "c:grubB done(HB_B prev)", 3: "c:rA ready", 4: "c:actA done", 5: "c:gemvB done", 6: "c:grubA done(HB_A)", 7: "c:rB ready", 8: "c:actB done",
         10: "p:idxA seen", 11: "p:A.r filled", 12: "p:A.z filled", 13: "p:A.h filled", 14: "p:idxB seen", 15: "p:B.r filled", 16: "p:B.z filled", 17: "p:B.h filled",
         27: "c:grubA X passed", 28: "c:grubA gemv done", 29: "c:grubA accb passed", 20: "s:HB_A seen", 21: "s:A sampled", 22: "s:A idx out", 24: "s:HB_B seen", 25: "s:B sampled", 26: "s:B idx out"}
for it in range(2, 5):
    ev = sorted((t[it, e] - t0, names[e]) for e in names if t[it, e])
    print("--- sample", 200 + it, "(cycles since sample 200 start; step = %d)" % (tw[it, 0, 0] - tw[it - 1, 0, 0]))
    base_t = (tw[it, 0, 0] - t0) if t[it, 0] == 0 else (t[it, 0] - t0)
    for c, nm in ev:
        if c - base_t > -30000: print("%8d  %s" % (c - base_t, nm))

# per-warp stamps of the compute warps (cycles since warp 0's start of the sample)
wn = ["start", "gemvA", "grubB", "rA rdy", "actA", "gemvB", "grubA", "rB rdy", "actB"]
for it in range(2, 4):
    print("--- sample", 200 + it, "per compute warp:", " ".join("%7s" % x for x in wn))
    for w in range(16):
        print("   warp %2d (sched %d)            " % (w, w % 4) + " ".join("%7d" % (tw[it, w, e] - tw[it, 0, 0]) for e in range(9)))
