#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu_final2.json 2> gpurun_out/r2_bench_2gpu_final2.err
echo rc=$?
tail -3 gpurun_out/r2_bench_2gpu_final2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_2gpu_final2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('parity_check',{}).get('ok_all_ranks'))
print('agg', d.get('embed_aggregate_resnet18'))
print('strong', d.get('strong_n100k'))
PY
