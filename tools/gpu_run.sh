mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.err; echo "bench8 rc=$?"
timeout 400 python bench.py --cpu-seconds 3 > gpurun_out/r2_bench_1gpu_final2.json 2> gpurun_out/r2_bench_1gpu_final2.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_8gpu_final','r2_bench_1gpu_final2'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value', d['value'], 'ms', d['ms_per_step'], '\n strong', d['strong_n100k']['value'], d['strong_n100k']['ms_per_step'], d['strong_n100k']['workload'][:40])
    except Exception as e: print(f,'no json', e, open(f'gpurun_out/{f}.err').read()[-800:])
PY
