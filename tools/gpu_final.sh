#!/bin/bash
# One-GPU profile session of the round's final state: all parity tests, smoke, bench lines of the three workloads + the CPU arm, batch-size
# probe, ncu launch list + full capture of the per-sample kernel.   usage: tools/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
make -C oracle port >/dev/null
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_${TAG}.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -c "import lpcnet_b200, json; print(json.dumps(lpcnet_b200.measure_smem_peak(0)))" | tee gpurun_out/smem_peak_${TAG}.json
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; tail -c 400 gpurun_out/bench_ref_${TAG}.json
for W in config2_float config5_decode; do
  timeout 500 python bench.py --gpus 1 --steps 10 --warmup 3 --frames 100 --workload $W > gpurun_out/bench_${W}_${TAG}.json 2>gpurun_out/bench_${W}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${W}_${TAG}.json").read().strip().splitlines()[-1])
print("$W", {k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "roofline frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("one_core_value"))
PY
done
timeout 300 python tools/probe_bench.py 14 256 1024 4096 4736 2>&1 | tail -4 | tee gpurun_out/probe_${TAG}.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launch_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpcnet_sample -s 4 -c 1 -o gpurun_out/prof_${TAG} -f \
    python bench.py --gpus 1 --steps 1 --warmup 3 --frames 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log
ls gpurun_out | head -40
