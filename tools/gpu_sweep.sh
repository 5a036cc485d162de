#!/bin/bash
# Variant sweep on one GPU: for every lpcnet_b200/variants/lib_<name>.so a bit-exactness check (golden + oracle tests through
# LPCNET_B200_SO) and the per-sample kernel time at 4096 streams; a timeline of the trace build; then the full parity suite and a
# bench line of the default build.  Every step has its own short timeout.     usage: tools/gpu_sweep.sh <tag>
TAG=${1:-sweep}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
make -C oracle port >/dev/null
for so in lpcnet_b200/variants/lib_*.so; do
  k=$(basename $so .so); k=${k#lib_}
  if [ "$k" = trace ]; then continue; fi
  par=$(LPCNET_B200_SO=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or ragged" 2>&1 | tail -1)
  t=$(LPCNET_B200_SO=$PWD/$so timeout 120 python tools/probe_bench.py 14 4096 2>&1 | tail -1)
  echo "$k | $par | $t" | tee -a gpurun_out/sweep_${TAG}.txt
done
if [ -f lpcnet_b200/variants/lib_trace.so ]; then
  LPCNET_B200_SO=$PWD/lpcnet_b200/variants/lib_trace.so timeout 120 python tools/trace_run.py > gpurun_out/trace_${TAG}.txt 2>&1; tail -64 gpurun_out/trace_${TAG}.txt
fi
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_${TAG}.txt
timeout 150 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "kernel ms", d["roofline"]["kernel_ms_per_launch"])
PY
tail -3 gpurun_out/bench_${TAG}.err
