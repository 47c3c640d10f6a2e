mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_4.log
tail -8 gpurun_out/r2_pytest_gpu_4.log
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_c.json 2> gpurun_out/r2_bench_pair_c.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_pair_c.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
except Exception as e: print('no bench json', e)
PY
timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_b.txt 2>&1; cat gpurun_out/r2_ptrace_b.txt
