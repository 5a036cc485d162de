#!/bin/bash
# compute-sanitizer memcheck + racecheck on a tiny run of both per-sample kernels (int8 and float flavours)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np, helpers as H, lpcnet_b200
from fixtures import make_feature_batch
for kind in ("int8", "float"):
    f = make_feature_batch(range(40), 4)
    b = lpcnet_b200.Batch(40, H.blob(kind), lpc_gamma=H.LPC_GAMMA)
    got = b.synthesize(f, samples_per_frame=24)           # 2 active frames x 24 samples: enough to exercise every phase
    print(kind, "ok", int(np.abs(got).max()))
    b.close()
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -6 gpurun_out/sanitizer_$tool.txt
done
