#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_zz_jpeg_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r2_jpeg_pytest2.txt
cat gpurun_out/r2_jpeg_pytest2.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
