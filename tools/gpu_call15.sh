mkdir -p gpurun_out
DSMIL_B200_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_9.log 2>&1; echo "pytest(pair) rc=$?" >> gpurun_out/r2_pytest_gpu_9.log
tail -3 gpurun_out/r2_pytest_gpu_9.log
DSMIL_B200_PAIR=1 timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_g.json 2> gpurun_out/r2_bench_pair_g.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_pair_g',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-800:])
PY
DSMIL_B200_PAIR=1 timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_g.txt 2>&1; head -30 gpurun_out/r2_ptrace_g.txt | cut -c1-150; grep -n "steady state" gpurun_out/r2_ptrace_g.txt
