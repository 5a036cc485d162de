#!/usr/bin/env python3
"""Per-stream digests of the UNTOUCHED reference (oracle/_ref, `make -C oracle ref`) at the sizes BASELINE.json names,
with DISTINCT inputs per stream (features seed 1000+s, packets seed 2000+s), plus the clamp fixture.

Run in the build container only (needs /root/reference); takes a few minutes on 8 cores.

  at_size_digests.npz
     config3_int8   uint64[4096]  first 8 bytes of sha256(pcm[s]) : 4096 streams x 100 frames, int8 build A
     config2_float  uint64[256]   256 streams x 1000 frames (10 s), float build B
     config5_decode uint64[1024]  1024 streams x 250 packets (10 s), int8 build A, lpcnet_decode
  clamp_A.npz       pcm[4][40*160] of model_int8_clamp.bin (sampling tree biased to large excitation: the output clips at
                    both +32767 and -32767, lpcnet.c:265-269)
"""
import json, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import helpers as H
from fixtures import make_feature_batch, make_packets

SIZES = {"config3_int8": (4096, 100), "config2_float": (256, 1000), "config5_decode": (1024, 250)}

if __name__ == "__main__":
    ONLY_CLAMP = "--clamp-only" in sys.argv
    out = {}
    if not ONLY_CLAMP:
        t0 = time.time()
        n, T = SIZES["config3_int8"]
        out["config3_int8"] = H.stream_digests(H.ref_synth(make_feature_batch(range(n), T), "A"))
        print("config3 done", time.time() - t0, flush=True)
        n, T = SIZES["config2_float"]
        out["config2_float"] = H.stream_digests(H.ref_synth(make_feature_batch(range(n), T), "B"))
        print("config2 done", time.time() - t0, flush=True)
        n, P = SIZES["config5_decode"]
        out["config5_decode"] = H.stream_digests(H.ref_decode(np.stack([make_packets(s, P) for s in range(n)]), "A"))
        print("config5 done", time.time() - t0, flush=True)
        np.savez_compressed(os.path.join(HERE, "at_size_digests.npz"), **out)
    pcm = H.ref_synth(make_feature_batch(range(4), 40), "A", kind="int8_clamp")
    clipped = [int((pcm == 32767).sum()), int((pcm == -32767).sum())]
    assert min(clipped) > 100, clipped
    np.savez_compressed(os.path.join(HERE, "clamp_A.npz"), pcm=pcm)
    cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    meta = {"cpu": cpu, "sizes": SIZES, "clamp_clipped_samples": clipped,
            "model_int8_clamp_sha256": __import__("hashlib").sha256(H.blob("int8_clamp")).hexdigest()}
    json.dump(meta, open(os.path.join(HERE, "at_size_meta.json"), "w"), indent=1)
    print(json.dumps(meta))
