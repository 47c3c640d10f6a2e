"""Evaluation epoch and classic-MIL drivers on the device: the batched `forward_bags` route of
`feed.eval_epoch` equals the one-bag-at-a-time reference loop shape (train_tcga.py:85-107), and a classic-MIL
store trains.  Written after this round's GPU budget was spent; runs after the parity suites on purpose."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,C", [(512, 2), (166, 1)])
def test_eval_epoch_batched_equals_per_bag_loop(D, C):
    import dsmil as mil
    from dsmil_wsi_b200.feed import DeviceBagStore, eval_epoch
    torch.manual_seed(D)
    net = mil.MILNet(mil.FCLayer(D, C), mil.BClassifier(D, C)).cuda()
    store = DeviceBagStore(D)
    for i in range(9):
        label = torch.zeros(C)
        label[i % C] = float(i % 2)
        store.add_bag(torch.rand(40 + 61 * i, D), label)
    crit = torch.nn.BCEWithLogitsLoss()
    loss_b, labels_b, preds_b = eval_epoch(net, store, crit, average=True, bags_per_launch=4)
    loss_1, labels_1, preds_1 = eval_epoch(net, store, crit, average=True, bags_per_launch=1)
    assert np.array_equal(labels_b, labels_1) and labels_b.shape == (9, C)
    assert np.allclose(preds_b, preds_1, atol=2e-6) and abs(loss_b - loss_1) < 2e-6
    # and against the reference's loop written out (train_tcga.py:91-106)
    total = 0.0
    with torch.no_grad():
        for feats, label in store.bags:
            ins, bag, _, _ = net(feats)
            mx, _ = torch.max(ins, 0)
            total += (0.5 * crit(bag.view(1, -1), label.view(1, -1)) + 0.5 * crit(mx.view(1, -1), label.view(1, -1))).item()
    assert abs(loss_b - total / len(store)) < 2e-6


def test_classic_mil_drivers_learn_on_the_device(tmp_path):
    import dsmil as mil
    from dsmil_wsi_b200 import formats
    from dsmil_wsi_b200.feed import (compute_pos_weight, cross_validation_set, mil_epoch_test, mil_epoch_train,
                                     mil_store)
    rng = np.random.default_rng(4)
    bags = [(int(i % 2), (rng.standard_normal((int(rng.integers(3, 12)), 166)) * 0.5 + 0.8 * (i % 2)).astype(np.float32))
            for i in range(40)]
    p = str(tmp_path / "musk_like.svm")
    formats.write_mil_svm(p, bags)
    all_bags = formats.mil_bags(formats.read_mil_svm(p), num_feats=166)
    train, test = cross_validation_set(all_bags, 4, 1)
    assert len(train) == 30 and len(test) == 10
    torch.manual_seed(0)
    net = mil.MILNet(mil.FCLayer(166, 1), mil.BClassifier(166, 1)).cuda()
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(compute_pos_weight(train)).cuda())
    opt = torch.optim.Adam(net.parameters(), lr=2e-3, betas=(0.5, 0.9), weight_decay=5e-3)
    tr_store, te_store = mil_store(train), mil_store(test)
    first = mil_epoch_test(net, te_store, crit)[0]
    for _ in range(8):
        mil_epoch_train(net, tr_store, crit, opt, order=range(len(tr_store)))
    loss, labels, preds = mil_epoch_test(net, te_store, crit)
    assert loss < 0.8 * first, (first, loss)
    acc = np.mean([(q > 0.5) == bool(l) for q, l in zip(preds, labels)])
    assert acc >= 0.8, acc
