#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_shard_backward_gpu.py tests/test_feed.py -x -q 2>&1 | tail -8 > gpurun_out/r2_bwd_pytest.txt
cat gpurun_out/r2_bwd_pytest.txt
timeout 200 python tools/train_step_probe.py > gpurun_out/r2_train_probe2.json 2>gpurun_out/r2_train_probe.err
cat gpurun_out/r2_train_probe2.json; tail -2 gpurun_out/r2_train_probe.err
PROBE_STEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_train_launches2.csv python tools/train_step_probe.py > gpurun_out/r2_train_ncu.log 2>&1
tail -1 gpurun_out/r2_train_ncu.log
