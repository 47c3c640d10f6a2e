mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_zz_acceptance_gpu.py > gpurun_out/r2_pytest_gpu_11.log 2>&1; echo "pytest(default) rc=$?" >> gpurun_out/r2_pytest_gpu_11.log
tail -4 gpurun_out/r2_pytest_gpu_11.log
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err; echo "bench rc=$?"
DSMIL_B200_PAIR=1 timeout 300 python bench.py --no-extras --cpu-seconds 1 > gpurun_out/r2_bench_pair_h.json 2> /dev/null
python - <<'PY'
import json
for f in ('r2_bench_1gpu_b','r2_bench_pair_h'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
        ex=d.get('extras') or {}
        for k,v in ex.items(): print('   ',k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','api')})
    except Exception as e: print(f,'no bench json', e)
PY
