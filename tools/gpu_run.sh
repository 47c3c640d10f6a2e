#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_zz_jpeg_gpu.py tests/test_zz_tree_gpu.py -x -q -k "embed_bag or tree" 2>&1 | tail -4 > gpurun_out/r2_jpeg_embed_pytest.txt
cat gpurun_out/r2_jpeg_embed_pytest.txt
timeout 600 python - > gpurun_out/r2_files_leg.json 2> gpurun_out/r2_files_leg.err <<'PY'
import json, sys, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
refmod = bench.load_reference_module()
print(json.dumps({"embed_from_files": bench.files_leg(dev, refmod)}, indent=1))
PY
grep -v "Computed" gpurun_out/r2_files_leg.err | tail -5 | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_files_leg.json'))
print(json.dumps(d["embed_from_files"]["compute_feats"], indent=1))
PY
