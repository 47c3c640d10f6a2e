mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.log
tail -15 gpurun_out/r2_pytest_gpu_2.log
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu_a.json 2> gpurun_out/r2_bench_1gpu_a.err; echo "bench1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu_a.json 2> gpurun_out/r2_bench_2gpu_a.err; echo "bench2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; echo "ref rc=$?"
tail -c 1500 gpurun_out/r2_bench_1gpu_a.err; tail -c 800 gpurun_out/r2_bench_2gpu_a.err
