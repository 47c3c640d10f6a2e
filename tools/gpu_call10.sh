mkdir -p gpurun_out
DSMIL_B200_PAIR=1 timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_c.txt 2>&1; tail -75 gpurun_out/r2_ptrace_c.txt
