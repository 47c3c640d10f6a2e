"""CTA-0 timeline of k_qmlp_sm100 (clock64 stamps via dsmil_debug_set_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import Weights, make_net
from dsmil_wsi_b200 import _lib
from dsmil_wsi_b200.sharded import CudaShardOps, milnet_params

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = torch.device("cuda", 0)
net = make_net(Weights(0), dev)
lib = _lib.load()
ops = CudaShardOps(milnet_params(net))
x = torch.rand(N, 512, device=dev)
for _ in range(2):
    ops.phase1(x, 0)
buf = torch.zeros(3 * 8 * 64, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
lib.dsmil_debug_set_trace(buf.data_ptr())
ops.phase1(x, 0)
torch.cuda.synchronize()
lib.dsmil_debug_set_trace(None)
t = buf.cpu().numpy().reshape(3, 8, 64)
t0 = t[2, 7, 0]
rel = lambda v: int(v - t0) if v else None
print("converter (third tile): per chunk  iter_start, stage_free, data_landed, published  [cycles since prologue]")
for kc in range(8):
    print(kc, [rel(t[0, e, kc]) for e in range(4)])
print("mma (third tile): per chunk  A_FULL, W_FULL, issued+commit")
for kc in range(8):
    print(kc, [rel(t[1, e, kc]) for e in range(3)])
print("epilogue: per tile  H1_FULL, A2_EMPTY ok, h1-epilogue done, q-epilogue(prev) done")
for it in range(8):
    print(it, [rel(t[2, e, it]) for e in (0, 3, 1, 2)])
print("mma L2(j): reached, A2_FULL ok, Q_EMPTY ok")
for it in range(8):
    print(it, [rel(t[1, e, it]) for e in (3, 4, 5)])
