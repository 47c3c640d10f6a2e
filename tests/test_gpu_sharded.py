"""Row-sharded forward on the GPU: G logical shards on one device (collectives replaced by local
concatenation, same kernels) must reproduce the single-call forward and the fp64 oracle; with >= 2
visible GPUs the real NCCL path is run as well."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_to_max
from helpers import build_net
from oracle import dsmil_oracle as orc

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name,G", [("shipped_tcga", 2), ("shipped_tcga", 8), ("musk_d166_n7", 8), ("musk_d166_n1", 2),
                                    ("pv_d96_c2", 3), ("lin_d512_c3", 4), ("tree_d1024_c2", 7)])
def test_virtual_shards_match_single_device_and_oracle(name, G):
    from dsmil_wsi_b200.sharded import CudaShardOps, milnet_params, virtual_sharded_forward
    g, p, X = load_golden(name)
    net = build_net(p).eval()
    x = torch.from_numpy(X).cuda()
    with torch.no_grad():
        c1, p1, A1, B1 = net(x)
    ops = CudaShardOps(milnet_params(net))
    c2, p2, A2, B2, crit = virtual_sharded_forward(ops, x, G)
    assert np.array_equal(_np(crit), g["idx"])
    assert torch.equal(c1, c2)
    assert rel_to_max(_np(A2), _np(A1)) < 2e-6 and rel_to_max(_np(B2), _np(B1)) < 2e-6
    assert rel_to_max(_np(p2), _np(p1)) < 1e-5
    t = orc.forward(X, p)
    assert rel_to_max(_np(A2), t.A) < 2e-5 and rel_to_max(_np(B2), t.B) < 1e-5


def test_giant_bag_shard_count_invariance():
    """BASELINE config 5 shape: N=100000 x 512, C=2; G in {1,2,4,8} give the same bag."""
    from dsmil_wsi_b200.sharded import CudaShardOps, milnet_params, virtual_sharded_forward
    p = orc.random_params(512, 2, 901, scale=2.0)
    X = orc.synthetic_bag(100000, 512, 902, "uniform")
    net = build_net(p).eval()
    x = torch.from_numpy(X).cuda()
    ops = CudaShardOps(milnet_params(net))
    ref = virtual_sharded_forward(ops, x, 1)
    t = orc.forward(X, p)
    assert np.array_equal(_np(ref[4]), t.idx)
    for G in (2, 4, 8):
        out = virtual_sharded_forward(ops, x, G)
        assert torch.equal(out[4], ref[4]) and torch.equal(out[0], ref[0])
        assert rel_to_max(_np(out[2]), _np(ref[2])) < 2e-6 and rel_to_max(_np(out[3]), _np(ref[3])) < 2e-6
    assert rel_to_max(_np(ref[2]), t.A) < 2e-5 and rel_to_max(_np(ref[3]), t.B) < 1e-5


def _nccl_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dsmil_wsi_b200.sharded import CudaShardOps, milnet_params, shard_bounds, sharded_forward
        g, p, X = load_golden("shipped_tcga")
        net = build_net(p, device=f"cuda:{rank}").eval()
        lo, hi = shard_bounds(X.shape[0], world)[rank]
        ops = CudaShardOps(milnet_params(net))
        c, pr, A, B, crit = sharded_forward(ops, torch.from_numpy(X[lo:hi]).cuda(), lo)
        torch.cuda.synchronize()
        ret[rank] = dict(lo=lo, hi=hi, A=_np(A), B=_np(B), pred=_np(pr), crit=_np(crit))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_nccl_two_ranks():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(2, port, ret), nprocs=2, join=True)
    g, p, X = load_golden("shipped_tcga")
    t = orc.forward(X, p)
    for r in range(2):
        o = ret[r]
        assert np.array_equal(o["crit"], t.idx)
        assert rel_to_max(o["A"], t.A[o["lo"]:o["hi"]]) < 2e-5 and rel_to_max(o["B"], t.B) < 1e-5
