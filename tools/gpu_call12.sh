mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_zz_feed_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu_5.log 2>&1; echo "pytest(fuse default) rc=$?" >> gpurun_out/r2_pytest_gpu_5.log
tail -12 gpurun_out/r2_pytest_gpu_5.log
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_fuse_a.json 2> gpurun_out/r2_bench_fuse_a.err; echo "bench fuse rc=$?"
DSMIL_B200_PAIR=1 timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_d.json 2> gpurun_out/r2_bench_pair_d.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_fuse_a','r2_bench_pair_d'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-800:])
PY
DSMIL_B200_PAIR=1 timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_d.txt 2>&1; head -60 gpurun_out/r2_ptrace_d.txt
