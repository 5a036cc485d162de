#!/usr/bin/env python3
"""Build kernel-geometry variants locally (tools/sweep.py build) and time them on the GPU (tools/sweep.py run)."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
VARIANTS = {
    "c16p7b3": ["LPCNET_NWC=16", "LPCNET_NWP=7", "LPCNET_GB=3"],
    "c16p7b4": ["LPCNET_NWC=16", "LPCNET_NWP=7", "LPCNET_GB=4"],
    "c16p7b2": ["LPCNET_NWC=16", "LPCNET_NWP=7", "LPCNET_GB=2"],
    "c16p3b5": ["LPCNET_NWC=16", "LPCNET_NWP=3", "LPCNET_GB=5"],
    "c12p3b6": ["LPCNET_NWC=12", "LPCNET_NWP=3", "LPCNET_GB=6"],
    "c12p7b4": ["LPCNET_NWC=12", "LPCNET_NWP=7", "LPCNET_GB=4"],
    "c24p7b2": ["LPCNET_NWC=24", "LPCNET_NWP=7", "LPCNET_GB=2"],
    "c24p3b3": ["LPCNET_NWC=24", "LPCNET_NWP=3", "LPCNET_GB=3"],
}
VARIANTS = {"c16p7": ["LPCNET_NWC=16","LPCNET_NWP=7"], "c24p7": ["LPCNET_NWC=24","LPCNET_NWP=7"], "c24p5": ["LPCNET_NWC=24","LPCNET_NWP=5"], "c12p7": ["LPCNET_NWC=12","LPCNET_NWP=7"], "c12p3": ["LPCNET_NWC=12","LPCNET_NWP=3"]}
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
