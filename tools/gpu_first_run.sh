#!/bin/bash
# First contact with the B200: build check, parity tests, then a tiny timing probe.  Everything under `timeout`.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -c "import lpcnet_b200; print('devices', lpcnet_b200.device_count())"
make -C oracle port >/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
timeout 300 python tools/probe_bench.py 18 2>&1 | tail -8
