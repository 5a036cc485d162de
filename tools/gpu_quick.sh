#!/bin/bash
# quick check: parity tests + bench line (no profiler)
TAG=${1:-q}
mkdir -p gpurun_out
make -C oracle port >/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["frac_of_smem_peak"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
tail -3 gpurun_out/bench_${TAG}.err
