mkdir -p gpurun_out
timeout 300 python bench.py --no-extras > gpurun_out/r2_bench_pair_b.json 2> gpurun_out/r2_bench_pair_b.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_pair_b.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
except Exception as e: print('no bench json', e)
PY
timeout 200 python tools/ptrace.py > gpurun_out/r2_ptrace_a.txt 2>&1; cat gpurun_out/r2_ptrace_a.txt
