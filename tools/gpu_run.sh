mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "graph or generic_route or nccl" > gpurun_out/r2_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_new.log
tail -6 gpurun_out/r2_pytest_new.log
