"""Host logic of the caller-side feed: fold bookkeeping pinned to train_mil.py's own functions
(tests/golden/formats/expected.npz), and the evaluation epoch checked on CPU against the reference's loop shape
(train_tcga.py:85-107 / train_mil.py:61-80) with a plain-torch stand-in for the operator (the product operator
needs a GPU: tests/test_feed.py, tests/test_zz_feed_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from dsmil_wsi_b200 import feed

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formats")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(FIX, "expected.npz")))


class TorchMIL(nn.Module):
    """dsmil.py:46-62 in plain torch (test stand-in; no forward_bags unless asked)."""

    def __init__(self, D, C, batched=False):
        super().__init__()
        self.fc = nn.Linear(D, C)
        self.q = nn.Sequential(nn.Linear(D, 128), nn.ReLU(), nn.Linear(128, 128), nn.Tanh())
        self.fcc = nn.Conv1d(C, C, kernel_size=D)
        self.calls = []
        if batched:
            self.forward_bags = lambda xs: (self.calls.append(len(xs)), [self(x) for x in xs])[1]

    def forward(self, x):
        c = self.fc(x)
        Q = self.q(x)
        _, idx = torch.sort(c, 0, descending=True)
        qm = self.q(x.index_select(0, idx[0]))
        A = torch.softmax(Q @ qm.t() / torch.sqrt(torch.tensor(128.0)), 0)
        B = (A.t() @ x).unsqueeze(0)
        return c, self.fcc(B).view(1, -1), A, B


def make_store(C, n=7, D=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    store = feed.DeviceBagStore(D, device="cpu")
    for i in range(n):
        label = torch.zeros(C)
        if C == 1:
            label[0] = float(i % 2)
        elif i % (C + 1) < C:
            label[i % (C + 1)] = 1
        store.add_bag(torch.randn(3 + 2 * i, D, generator=g), label)
    return store


@pytest.mark.parametrize("C,average,batched", [(1, False, False), (2, False, True), (2, True, True), (3, True, False)])
def test_eval_epoch_matches_reference_loop_shape(C, average, batched):
    torch.manual_seed(C)
    net = TorchMIL(12, C, batched=batched)
    store = make_store(C)
    crit = nn.BCEWithLogitsLoss()
    loss, labels, preds = feed.eval_epoch(net, store, crit, average=average, bags_per_launch=3)
    # the reference's loop (train_tcga.py:91-106), one bag at a time with .item() per bag
    total, ref_labels, ref_preds = 0.0, [], []
    with torch.no_grad():
        for feats, label in store.bags:
            ins, bag, _, _ = net(feats)
            mx, _ = torch.max(ins, 0)
            l = 0.5 * crit(bag.view(1, -1), label.view(1, -1)) + 0.5 * crit(mx.view(1, -1), label.view(1, -1))
            total += l.item()
            ref_labels.append(label.squeeze().numpy().astype(int))
            ref_preds.append(((torch.sigmoid(mx) + torch.sigmoid(bag)) if average else torch.sigmoid(bag)).squeeze().numpy())
    assert abs(loss - total / len(store)) < 1e-6
    assert labels.shape == preds.shape == (len(store), C) and labels.dtype.kind == "i"
    assert np.array_equal(labels.reshape(len(store), -1), np.array(ref_labels).reshape(len(store), -1))
    assert np.allclose(preds.reshape(len(store), -1), np.array(ref_preds).reshape(len(store), -1), atol=1e-7)
    if batched:
        assert net.calls == [3, 3, 1]            # groups of bags_per_launch through forward_bags
    assert not net.training                      # milnet.eval() as train_tcga.py:86


def test_eval_epoch_rejects_an_empty_store():
    with pytest.raises(ValueError, match="empty"):
        feed.eval_epoch(TorchMIL(4, 1), feed.DeviceBagStore(4, device="cpu"), nn.BCEWithLogitsLoss())


def test_fold_bookkeeping_is_the_reference(gold):
    for n_items, fold in ((23, 5), (92, 10), (10, 10)):
        for index in range(fold):
            tr, te = feed.cross_validation_set(list(range(n_items)), fold, index)
            assert np.array_equal(tr, gold[f"cv_{n_items}_{fold}_{index}_train"])
            assert np.array_equal(te, gold[f"cv_{n_items}_{fold}_{index}_test"])
    with pytest.raises(ValueError, match="folds"):
        feed.cross_validation_set([1, 2, 3], 5, 0)
    labs = gold["pos_weight_labels"].tolist()
    assert feed.compute_pos_weight([(l, None) for l in labs]) == float(gold["pos_weight"])
    with pytest.raises(ZeroDivisionError):
        feed.compute_pos_weight([(0, None), (-1, None)])


def test_classic_mil_store_and_test_epoch(tmp_path):
    from dsmil_wsi_b200 import formats
    rng = np.random.default_rng(2)
    bags = [(int(i % 2), (rng.standard_normal((int(rng.integers(2, 9)), 6)) + 2.0 * (i % 2)).astype(np.float32))
            for i in range(10)]
    p = str(tmp_path / "toy.svm")
    formats.write_mil_svm(p, bags)
    store = feed.mil_store(formats.mil_bags(formats.read_mil_svm(p), num_feats=6), device="cpu")
    assert len(store) == 10 and store.D == 6
    for (label, x), (f, l) in zip(bags, store.bags):
        assert l.shape == (1, 1) and float(l) == label and np.array_equal(f.numpy(), x)
    torch.manual_seed(0)
    net = TorchMIL(6, 1)
    loss, labels, preds = feed.mil_epoch_test(net, store, nn.BCEWithLogitsLoss(), bags_per_launch=4)
    assert labels == [b[0] for b in bags] and len(preds) == 10 and all(np.ndim(q) == 0 and 0 < q < 1 for q in preds)
    with torch.no_grad():
        ref = np.mean([float(0.5 * nn.BCEWithLogitsLoss()(net(f)[1].view(1, -1), l) +
                             0.5 * nn.BCEWithLogitsLoss()(net(f)[0].max(0)[0].view(1, -1), l)) for f, l in store.bags])
    assert abs(loss - ref) < 1e-6
    with pytest.raises(ValueError):
        feed.mil_store([])
