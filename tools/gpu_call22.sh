mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu_d.json 2> gpurun_out/r2_bench_2gpu_d.err; echo "bench2 graph rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_2gpu_d',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], d['config'].get('step_launch'), '\n strong', d.get('strong_n100k',{}).get('value'), d.get('strong_n100k',{}).get('ms_per_step'), '\n breakdown', d.get('step_breakdown'), '\n parity', d.get('parity_check',{}).get('ok_all_ranks'))
    except Exception as e: print(f,'no bench json', e, open(f'gpurun_out/{f}.err').read()[-1500:])
PY
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_zz_shard_backward_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
