#!/bin/bash
# last check of a HEAD: all GPU parity tests, smoke, one bench line, the single-stream latency probe.   usage: tools/gpu_verify.sh <tag>
TAG=${1:-v}
mkdir -p gpurun_out
make -C oracle port >/dev/null
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_${TAG}.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "kernel ms", d["roofline"]["kernel_ms_per_launch"], "frac", d["roofline"]["frac"], d["roofline"]["traffic_source"])
PY
timeout 120 python tools/probe_single.py 2>&1 | tail -2 | tee gpurun_out/probe_single_${TAG}.txt
