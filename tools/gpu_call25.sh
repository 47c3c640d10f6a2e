mkdir -p gpurun_out
timeout 200 python tools/ktrace.py 100000 > gpurun_out/r2_ktrace_old.txt 2>&1; cat gpurun_out/r2_ktrace_old.txt
