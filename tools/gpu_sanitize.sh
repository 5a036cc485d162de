#!/bin/bash
# compute-sanitizer memcheck + racecheck on a tiny run of both per-sample kernels (int8 and float flavours)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np, helpers as H, lpcnet_b200
from fixtures import make_feature_batch
for kind, spc in (("int8", "32"), ("int8", "17"), ("int8", "5"), ("float", "32"), ("float", "2"), ("float", "3")):
    os.environ["LPCNET_B200_STREAMS_PER_CTA"] = spc     # int8: 32 = full CTA + ragged one, 17 = two halves with dead slots, 5 = half-A-only schedule; float: 32 = lane==stream kernel, 2/3 = neuron-per-lane kernel
    f = make_feature_batch(range(40), 4)
    b = lpcnet_b200.Batch(40, H.blob(kind), lpc_gamma=H.LPC_GAMMA)
    got = b.synthesize(f, samples_per_frame=24)           # 2 active frames x 24 samples: enough to exercise every phase
    print(kind, spc, "ok", int(np.abs(got).max()))
    b.close()
PY
# racecheck runs on the checking build (python tools/build_check_variant.py): see warp_arrive() in sample_kernel.cu
for tool in memcheck racecheck; do
  if [ $tool = racecheck ] && [ -f lpcnet_b200/variants/lib_arriveall.so ]; then export LPCNET_B200_SO=lpcnet_b200/variants/lib_arriveall.so; fi
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -6 gpurun_out/sanitizer_$tool.txt
done
