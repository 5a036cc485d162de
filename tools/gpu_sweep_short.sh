#!/bin/bash
# per variant library: bit-exactness subset + per-sample kernel time at 4096 streams (two runs).   usage: tools/gpu_sweep_short.sh <tag>
TAG=${1:-s}
mkdir -p gpurun_out
make -C oracle port >/dev/null
for so in lpcnet_b200/variants/lib_*.so; do
  k=$(basename $so .so); k=${k#lib_}
  par=$(LPCNET_B200_SO=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or ragged or grid_shapes" 2>&1 | tail -1)
  t1=$(LPCNET_B200_SO=$PWD/$so timeout 120 python tools/probe_bench.py 14 4096 2>&1 | tail -1 | sed 's/ (kernel).*//')
  t2=$(LPCNET_B200_SO=$PWD/$so timeout 120 python tools/probe_bench.py 14 4096 4736 2>&1 | tail -2 | sed 's/ (kernel).*//' | tr '\n' ';')
  echo "$k | $par | $t1 | $t2" | tee -a gpurun_out/sweep_${TAG}.txt
done
