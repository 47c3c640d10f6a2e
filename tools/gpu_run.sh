#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_pytest_gpu_final2.txt
cat gpurun_out/r2_pytest_gpu_final2.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_1gpu_final2.json 2> gpurun_out/r2_bench_1gpu_final2.err
tail -4 gpurun_out/r2_bench_1gpu_final2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_1gpu_final2.json') if l.startswith('{')][-1])
ex=d.get('extras',{})
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
print('train', ex.get('train_n15000_c1'))
print('embed', {k:v for k,v in ex.get('embed_resnet18_in',{}).items() if k in ('value','ms_per_batch','ms_per_batch_nchw','speedup','unavailable')})
ef=ex.get('embed_from_files',{})
print('files', ef.get('compute_feats', ef))
print('agg', d.get('embed_aggregate_resnet18'))
PY
