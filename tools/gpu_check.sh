#!/bin/bash
# Quick GPU check: a subset of the parity tests, one bench line, the launch list.  usage: tools/gpu_check.sh <tag> [pytest files...]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
FILES=${@:-tests/test_gpu_parity.py tests/test_gpu_variants.py}
timeout 1200 python -m pytest $FILES -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "kernel ms", d["roofline"]["kernel_ms_per_launch"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r)>5]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki].split("(")[0],[]).append(float(r[vi].replace(",","")))
    except ValueError: pass
for k,v in agg.items(): print(k,len(v),round(sum(v)/len(v)/1e3,1),"us")
PY
