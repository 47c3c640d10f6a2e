mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu_a.json 2> gpurun_out/r2_bench_8gpu_a.err; echo "bench8 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_8gpu_a',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['config'].get('step_launch'), '\n strong', d.get('strong_n100k',{}).get('value'), d.get('strong_n100k',{}).get('ms_per_step'), d.get('strong_n100k',{}).get('step_launch'), '\n breakdown', d.get('step_breakdown'), '\n parity', d.get('parity_check',{}).get('ok_all_ranks'), 'e2e', d.get('e2e',{}).get('value'))
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-1500:])
PY
