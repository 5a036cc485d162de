#!/bin/bash
mkdir -p gpurun_out
make -C oracle port >/dev/null
for v in 0,0 1,0 1,24 1,48 0,24; do
  t=$(LPCNET_B200_LPT=$v timeout 120 python tools/probe_bench.py 14 4096 4096 2>&1 | tail -2 | sed 's/ (kernel).*//' | tr '\n' ';')
  echo "LPT=$v | $t" | tee -a gpurun_out/lpt_r02v.txt
done
LPCNET_B200_SO=$PWD/lpcnet_b200/variants/lib_trace.so timeout 120 python tools/trace_run.py > gpurun_out/trace_r02v.txt 2>&1; tail -36 gpurun_out/trace_r02v.txt | head -20
