#!/bin/bash
# GPU parity tests only (+ smoke + one short bench line).  usage: tools/gpu_tests.sh <tag> [pytest -k expression]
TAG=${1:-t}
mkdir -p gpurun_out
make -C oracle port >/dev/null
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 2400 python -m pytest tests -m gpu -q "${K[@]}" 2>&1 | tail -40 | tee gpurun_out/pytest_${TAG}.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "kernel ms", d["roofline"]["kernel_ms_per_launch"])
PY
tail -3 gpurun_out/bench_${TAG}.err
