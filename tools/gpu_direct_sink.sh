#!/bin/bash
# 2-GPU experiment: PCM gather by DMA sink (default) vs direct peer stores from the per-sample kernel.  usage: gpurun --gpus 2 -- tools/gpu_direct_sink.sh
mkdir -p gpurun_out
for MODE in sink direct sink direct; do
  if [ $MODE = direct ]; then export LPCNET_B200_BENCH_DIRECT_SINK=1; else unset LPCNET_B200_BENCH_DIRECT_SINK; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_ds_$MODE.json 2> gpurun_out/bench_ds_$MODE.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ds_$MODE.json").read().strip().splitlines()[-1])
    print("$MODE", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d.get("pcm_gather",{}).get("verified"))
except Exception as e:
    print("$MODE failed", e); print(open("gpurun_out/bench_ds_$MODE.err").read()[-1500:])
PY
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
