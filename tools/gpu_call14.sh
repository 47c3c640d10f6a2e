mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_7.log 2>&1; echo "pytest(default) rc=$?" >> gpurun_out/r2_pytest_gpu_7.log
tail -4 gpurun_out/r2_pytest_gpu_7.log
DSMIL_B200_PAIR=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_8.log 2>&1; echo "pytest(pair) rc=$?" >> gpurun_out/r2_pytest_gpu_8.log
tail -4 gpurun_out/r2_pytest_gpu_8.log
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_old_b.json 2> gpurun_out/r2_bench_old_b.err; echo "bench old rc=$?"
DSMIL_B200_PAIR=1 timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_f.json 2> gpurun_out/r2_bench_pair_f.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_old_b','r2_bench_pair_f'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-800:])
PY
DSMIL_B200_PAIR=1 timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_f.txt 2>&1; head -62 gpurun_out/r2_ptrace_f.txt | cut -c1-150
