mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_zz_feed_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_12.log 2>&1; echo "pytest(default) rc=$?" >> gpurun_out/r2_pytest_gpu_12.log
tail -4 gpurun_out/r2_pytest_gpu_12.log
timeout 400 python bench.py --no-extras --cpu-seconds 2 > gpurun_out/r2_bench_1gpu_c.json 2> gpurun_out/r2_bench_1gpu_c.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_1gpu_c',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f,'no bench json', e)
PY
