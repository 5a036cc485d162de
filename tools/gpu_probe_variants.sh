#!/bin/bash
# timing only: per-sample kernel time at 4096 streams for every variant library (no parity check: experiment builds may be wrong on purpose)
mkdir -p gpurun_out
make -C oracle port >/dev/null
for so in lpcnet_b200/variants/lib_*.so; do
  k=$(basename $so .so); k=${k#lib_}
  t=$(LPCNET_B200_SO=$PWD/$so timeout 120 python tools/probe_bench.py 14 4096 2>&1 | tail -1)
  echo "$k | $t" | tee -a gpurun_out/probe_variants.txt
done
