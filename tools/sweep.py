#!/usr/bin/env python3
"""Build kernel-geometry variants locally (tools/sweep.py build) and time them on the GPU (tools/sweep.py run)."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
# geometry / option variants of the int8 per-sample kernel (compile-time macros of sample_kernel.cu / engine.h).  Measured in
# round 1: 16 compute warps + 6 producers is the best geometry (12 and 24 compute warps are within 5 %).
VARIANTS = {
    "default": [],                                   # first three tree levels evaluated side by side
    "tree_serial": ["LPCNET_TREE_PAR=0"],
}
if sys.argv[1] == "build":
    from lpcnet_b200 import build
    for k, v in VARIANTS.items():
        build.build_variant(k, v)
else:
    for k in VARIANTS:
        so = os.path.join(ROOT, "lpcnet_b200", "variants", "lib_%s.so" % k)
        if not os.path.exists(so):
            continue
        env = dict(os.environ, LPCNET_B200_SO=so)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_bench.py"), "14", "4096"], env=env, capture_output=True, text=True, timeout=300)
        print(k, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
