make -C oracle port >/dev/null
LPCNET_B200_SO=$PWD/lpcnet_b200/variants/lib_trace.so timeout 120 python tools/trace_run.py > gpurun_out/trace_r02t.txt 2>&1; tail -64 gpurun_out/trace_r02t.txt
