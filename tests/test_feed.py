"""Caller-side bag feed: device gather == torch indexing; an epoch over device-resident bags."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D", [(1000, 512), (37, 166), (5, 4)])
def test_gather_rows_is_indexing(N, D):
    from dsmil_wsi_b200.feed import dropout_patches, gather_rows
    x = torch.randn(N, D, device="cuda")
    idx = torch.randperm(N, device="cuda")[: max(1, N // 2)]
    assert torch.equal(gather_rows(x, idx), x[idx])
    g = torch.Generator(device="cuda").manual_seed(3)
    y = dropout_patches(x, 0.75, g)
    g = torch.Generator(device="cuda").manual_seed(3)
    ref = x[torch.randperm(N, device="cuda", generator=g)[: int(N * 0.75)]]
    assert y.shape == (int(N * 0.75), D) and torch.equal(y, ref)


def test_epoch_over_device_store_learns():
    import dsmil as mil
    from dsmil_wsi_b200.feed import DeviceBagStore, train_epoch
    torch.manual_seed(0)
    store = DeviceBagStore(64)
    for i in range(12):
        y = float(i % 2)
        feats = torch.randn(50 + 7 * i, 64) + (1.5 * y)
        store.add_stacked(torch.cat([feats, torch.full((feats.shape[0], 1), y)], 1))
    net = mil.MILNet(mil.FCLayer(64, 1), mil.BClassifier(64, 1)).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=2e-3, betas=(0.5, 0.9))
    crit = torch.nn.BCEWithLogitsLoss()
    losses = [train_epoch(net, store, crit, opt, dropout_patch=0.1, order=range(len(store))) for _ in range(6)]
    assert losses[-1] < 0.7 * losses[0], losses
