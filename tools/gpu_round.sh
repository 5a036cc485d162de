#!/bin/bash
# One GPU session: parity tests, smoke, benchmark, ncu launch list, ncu full capture of the per-sample kernel.
# usage: tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
make -C oracle port >/dev/null
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; tail -c 1500 gpurun_out/bench_ref_${TAG}.json
# every launch with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_launch_${TAG}.log
# the per-sample kernel, full set, source view
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lpcnet_sample -s 4 -c 1 -o gpurun_out/prof_${TAG} -f \
    python bench.py --gpus 1 --steps 1 --warmup 3 --frames 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out
for W in config2_float config5_decode; do
  timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_${TAG}.json 2>gpurun_out/bench_${W}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${W}_${TAG}.json").read().strip().splitlines()[-1])
print("$W", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["config"]["workload"][:60])
PY
done
# throughput vs batch size (device-resident, kernel-only and wall): the grid follows the SM count
timeout 600 python tools/probe_bench.py 14 256 1024 4096 4736 9472 2>&1 | tail -6 | tee gpurun_out/probe_${TAG}.txt
