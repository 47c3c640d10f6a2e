mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.log
timeout 120 tools/chainbench > gpurun_out/r2_chainbench.txt 2>&1
timeout 300 python bench.py > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
tail -3 gpurun_out/r2_pytest_gpu_1.log; cat gpurun_out/r2_bench_base.json | head -c 1500
