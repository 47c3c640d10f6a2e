mkdir -p gpurun_out
DSMIL_B200_PAIR=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_6.log 2>&1; echo "pytest(pair) rc=$?" >> gpurun_out/r2_pytest_gpu_6.log
tail -6 gpurun_out/r2_pytest_gpu_6.log
DSMIL_B200_PAIR=1 timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_e.json 2> gpurun_out/r2_bench_pair_e.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_pair_e',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-800:])
PY
DSMIL_B200_PAIR=1 timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_e.txt 2>&1; head -75 gpurun_out/r2_ptrace_e.txt | cut -c1-150
