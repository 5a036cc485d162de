"""Quick device-timing probe (not the benchmark): samples/s of the per-sample kernel for a few batch sizes."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import helpers as H
import lpcnet_b200
from fixtures import make_feature_batch

T = int(sys.argv[1]) if len(sys.argv) > 1 else 18
for n in [int(a) for a in sys.argv[2:]] or [32, 256, 1024, 4096, 4736]:
    f = make_feature_batch(range(n), T)          # distinct streams
    b = lpcnet_b200.Batch(n, H.blob("int8"), lpc_gamma=H.LPC_GAMMA)
    b.synthesize(f[:, :4])          # warm-up (also consumes the 2 silent frames)
    t0 = time.time()
    b.synthesize(f[:, 4:])
    wall = time.time() - t0
    ms, k = b.last_sample_kernel_ms()
    samples = n * (T - 4) * 160
    print("n=%5d frames=%d sample-kernel %.2f ms -> %.3e samples/s (kernel), wall %.1f ms -> %.3e samples/s e2e, launches %d"
          % (n, T - 4, ms, samples / (ms * 1e-3), wall * 1e3, samples / wall, k), flush=True)
    b.close()
