mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_zz_acceptance_gpu.py > gpurun_out/r2_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_3.log
tail -30 gpurun_out/r2_pytest_gpu_3.log
timeout 400 python bench.py --no-extras > gpurun_out/r2_bench_pair_a.json 2> gpurun_out/r2_bench_pair_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_pair_a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_pair_a.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
except Exception as e: print('no bench json', e)
PY
