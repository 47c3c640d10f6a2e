#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 300 python tools/exp_channels_last.py > gpurun_out/r2_exp_channels_last.json 2> gpurun_out/r2_exp_cl.err
tail -3 gpurun_out/r2_exp_cl.err; cat gpurun_out/r2_exp_channels_last.json
cat > /tmp/jl.py <<'PY'
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from dsmil_wsi_b200 import jpeg
dev = torch.device('cuda', 0)
files = bench.synth_patch_files(128, seed=5)
pb = jpeg.parse_batch(files, pin=True)
dec = jpeg.JpegBatchDecoder(dev)
x = torch.empty(128, 3, 224, 224, device=dev)
for _ in range(3):
    dec.decode(pb, out_f32=x)
torch.cuda.synchronize()
big = files * 8
pb2 = jpeg.parse_batch(big, pin=True)
x2 = torch.empty(1024, 3, 224, 224, device=dev)
print('ms 1024:', bench.cuda_time_ms(lambda: dec.decode(pb2, out_f32=x2), 5, warm=2))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_jpeg_launches.csv python /tmp/jl.py > gpurun_out/r2_jpeg_ncu.log 2>&1
tail -3 gpurun_out/r2_jpeg_ncu.log
grep -c k_jpeg gpurun_out/r2_jpeg_launches.csv
timeout 120 python /tmp/jl.py
