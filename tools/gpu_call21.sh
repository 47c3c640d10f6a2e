mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err; echo "bench2 graph rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 --no-graph --no-extras > gpurun_out/r2_bench_2gpu_c.json 2> gpurun_out/r2_bench_2gpu_c.err; echo "bench2 eager rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_2gpu_b','r2_bench_2gpu_c'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['config'].get('step_launch'), 'strong', d.get('strong_n100k'), 'parity', d.get('parity_check'))
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-1500:])
PY
