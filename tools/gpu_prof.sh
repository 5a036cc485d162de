#!/bin/bash
# ncu full capture of the per-sample kernel only (3 frames), for tools/seg_profile.py.  usage: tools/gpu_prof.sh <tag>
TAG=${1:-p}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lpcnet_sample -s 4 -c 1 -o gpurun_out/prof_${TAG} -f \
    python bench.py --gpus 1 --steps 1 --warmup 3 --frames 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
