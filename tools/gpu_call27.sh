mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_embedder_gpu.py tests/test_embed.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_embedder.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_embedder.log
tail -12 gpurun_out/r2_pytest_embedder.log
timeout 500 python bench.py --cpu-seconds 4 > gpurun_out/r2_bench_1gpu_d.json 2> gpurun_out/r2_bench_1gpu_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu_d.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], d['roofline']['per_kernel_ms'], 'frac', d['roofline']['frac'])
print('embed_agg', d.get('embed_aggregate_resnet18'))
print('embed', (d.get('extras') or {}).get('embed_resnet18_in'))
print('strong', d.get('strong_n100k'))
PY
tail -c 600 gpurun_out/r2_bench_1gpu_d.err
