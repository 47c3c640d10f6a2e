#!/bin/bash
# One gpurun call's worth of end-of-round validation:  /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_run.sh'
mkdir -p gpurun_out
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -2 gpurun_out/bench_1gpu.err | cut -c1-300
tail -c 1500 gpurun_out/bench_1gpu.json
