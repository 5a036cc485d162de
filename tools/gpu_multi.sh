#!/bin/bash
# Two-GPU session: all GPU parity tests (incl. tests/test_gpu_multi.py's >= 2 device cases), bench at N=1 and N=2 (torchrun).
# usage (from the build container): gpurun --gpus 2 -- tools/gpu_multi.sh <tag>
TAG=${1:-m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
make -C oracle port >/dev/null
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_${TAG}.txt
timeout 200 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
tail -3 gpurun_out/bench_${TAG}_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
tail -5 gpurun_out/bench_${TAG}_n2.err
python - <<PY
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "kernel ms", d["roofline"]["kernel_ms_per_launch"], d.get("pcm_gather"))
    except Exception as e:
        print(n, "no line:", e)
PY
