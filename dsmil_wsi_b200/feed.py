"""Caller-side bag feed for training (reference train_tcga.py:55-83).

The reference's training loop pays, per bag and per epoch: `torch.load(item, map_location='cuda:0')` (disk ->
GPU), a CPU `torch.randperm` + advanced-indexing gather for `dropout_patches`, and a `loss.item()` sync.  Once
the operator takes tens of microseconds those dominate.  Here the bags of the `.pt` cache format
(`[N, D + C]` = features || label repeated per row, train_tcga.py:36-51) live on the device, the patch-dropout
permutation is drawn on the device and applied by `dsmil_gather_rows` (our kernel), and the running loss stays
on the device until the epoch ends.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import functional as Fn


def gather_rows(feats: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[m] = feats[idx[m]] on the device (== `feats[idx]` of train_tcga.py:82)."""
    Fn.require_cuda(feats, "feats")
    feats = Fn._f32c(feats)
    idx = idx.to(device=feats.device, dtype=torch.int64).contiguous()
    M, D = int(idx.numel()), int(feats.shape[1])
    with torch.cuda.device(feats.device):
        out = torch.empty(M, D, dtype=torch.float32, device=feats.device)
        _lib.check(_lib.load().dsmil_gather_rows(feats.data_ptr(), int(feats.shape[0]), D, idx.data_ptr(), M,
                                                 out.data_ptr(), Fn._stream()), "dsmil_gather_rows")
    return out


def dropout_patches(feats: torch.Tensor, p: float, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """train_tcga.py:78-83 with the arguments of the call site (`dropout_patches(bag_feats, 1 - dropout_patch)`):
    keep int(N * p) rows in random order (p = 1 -> a full random permutation), everything on the device."""
    n = int(feats.shape[0])
    keep = int(n * p)
    perm = torch.randperm(n, device=feats.device, generator=generator)[:keep]
    return gather_rows(feats, perm)


class DeviceBagStore:
    """Bags resident in HBM (180 GB holds thousands of 15 000 x 512 bags), in the `.pt` cache layout."""

    def __init__(self, feats_size: int, device="cuda"):
        self.D = feats_size
        self.device = torch.device(device)
        self.bags: List[Tuple[torch.Tensor, torch.Tensor]] = []

    def add_stacked(self, stacked: torch.Tensor) -> None:
        """stacked: [N, D + C] (train_tcga.py:47-51)."""
        st = stacked.to(self.device, dtype=torch.float32)
        self.bags.append((st[:, : self.D].contiguous(), st[0, self.D:].clone().unsqueeze(0)))

    def add_files(self, paths: Iterable[str]) -> None:
        for p in paths:
            self.add_stacked(torch.load(p, map_location="cpu"))

    def add_bag(self, feats: torch.Tensor, label: torch.Tensor) -> None:
        """feats [N, D] and label [C] given separately (no stacked copy)."""
        if feats.dim() != 2 or feats.shape[1] != self.D:
            raise ValueError(f"feats must be [N, {self.D}], got {tuple(feats.shape)}")
        self.bags.append((feats.to(self.device, dtype=torch.float32, non_blocking=True).contiguous(),
                          label.to(self.device, dtype=torch.float32).reshape(1, -1)))

    def add_bins(self, paths: Iterable[str]) -> None:
        """Binary bag containers (formats.write_bag_bin): payload read straight into pinned memory, then one
        asynchronous H2D copy per bag -- no text parse, no [N, D + C] intermediate."""
        from . import formats
        pin = self.device.type == "cuda"
        for p in paths:
            feats, label = formats.read_bag_bin(p, pin_memory=pin)
            self.add_bag(feats, label)
        if pin:
            torch.cuda.current_stream(self.device).synchronize()   # pinned sources may be freed after this

    def add_index(self, index_csv: str, num_classes: int, tcga_default: bool = False, workers: int = 4) -> None:
        """The reference's route (train_tcga.py:245-250 + :36-51) without the temp_train/*.pt detour.  Bag CSVs
        are parsed by the native reader on `workers` threads (the C call releases the GIL), in index order."""
        from concurrent.futures import ThreadPoolExecutor
        from . import formats
        rows = formats.read_dataset_index(index_csv)
        paths = [formats.tcga_default_feats_path(e) if tcga_default else e for e, _ in rows]
        with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
            for feats, (_, label) in zip(pool.map(formats.read_bag_csv, paths), rows):
                self.add_bag(torch.from_numpy(feats), torch.from_numpy(formats.bag_label(label, num_classes)))

    def __len__(self):
        return len(self.bags)


def train_epoch(milnet, store: DeviceBagStore, criterion, optimizer, dropout_patch: float = 0.0,
                order: Optional[Sequence[int]] = None, generator: Optional[torch.Generator] = None) -> float:
    """One epoch of train_tcga.train() (train_tcga.py:55-76) over device-resident bags; one host sync per epoch."""
    milnet.train()
    total = torch.zeros((), device=store.device)
    order = list(order) if order is not None else torch.randperm(len(store)).tolist()
    for i in order:
        feats, label = store.bags[i]
        optimizer.zero_grad()
        x = dropout_patches(feats, 1 - dropout_patch, generator)
        ins_prediction, bag_prediction, _, _ = milnet(x)
        max_prediction, _ = torch.max(ins_prediction, 0)
        loss = 0.5 * criterion(bag_prediction.view(1, -1), label.view(1, -1)) + \
            0.5 * criterion(max_prediction.view(1, -1), label.view(1, -1))
        loss.backward()
        optimizer.step()
        total += loss.detach()
    return float(total.item()) / max(1, len(order))


def eval_epoch(milnet, store: DeviceBagStore, criterion, average: bool = False, bags_per_launch: int = 16,
               dropout_patch: float = 0.0, generator: Optional[torch.Generator] = None):
    """The measuring half of train_tcga.test() (train_tcga.py:85-107): mean loss, labels [n_bags, C] and
    sigmoid predictions [n_bags, C] (`average=True`: sigmoid(max) + sigmoid(bag), as `args.average`).  The ROC /
    threshold logic that follows in the reference (:108-132) is the caller's and works on these arrays.

    Bags go through `milnet.forward_bags` `bags_per_launch` at a time when the module has it (one batched launch
    sequence for the whole group, the slides/sec path), else one `milnet(x)` per bag; the loss terms stay on the
    device and there is one host sync per epoch instead of two `.item()` per bag.  The reference also routes test
    bags through `dropout_patches` (a row permutation when dropout_patch == 0); a permutation does not change
    any output of the operator except the order of the per-instance rows, so it is skipped unless rows are
    actually dropped."""
    milnet.eval()
    n = len(store)
    if n == 0:
        raise ValueError("eval_epoch over an empty store")
    losses, labels, preds = [], [], []
    batched = hasattr(milnet, "forward_bags") and bags_per_launch > 1
    with torch.no_grad():
        for s in range(0, n, max(1, bags_per_launch)):
            group = store.bags[s:s + max(1, bags_per_launch)]
            xs = [dropout_patches(f, 1 - dropout_patch, generator) if dropout_patch > 0 else f for f, _ in group]
            outs = milnet.forward_bags(xs) if batched else [milnet(x) for x in xs]
            for (ins_prediction, bag_prediction, _, _), (_, label) in zip(outs, group):
                max_prediction, _ = torch.max(ins_prediction, 0)
                losses.append(0.5 * criterion(bag_prediction.view(1, -1), label.view(1, -1)) +
                              0.5 * criterion(max_prediction.view(1, -1), label.view(1, -1)))
                labels.append(label.view(-1))
                p = torch.sigmoid(bag_prediction).view(-1)
                preds.append(torch.sigmoid(max_prediction).view(-1) + p if average else p)
    loss = float(torch.stack([l.reshape(()) for l in losses]).mean().item())
    return loss, torch.stack(labels).cpu().numpy().astype(int), torch.stack(preds).cpu().numpy()


# ---- classic-MIL drivers (train_mil.py) ---------------------------------------------------------------------


def mil_store(bags: Sequence[Tuple[int, "np.ndarray"]], device="cuda") -> DeviceBagStore:
    """`formats.mil_bags(...)` output -> device-resident store (label as a [1, 1] target, train_mil.py:49)."""
    if not bags:
        raise ValueError("no bags")
    store = DeviceBagStore(int(bags[0][1].shape[1]), device=device)
    for label, x in bags:
        store.add_bag(torch.as_tensor(x, dtype=torch.float32), torch.tensor([float(label)]))
    return store


def cross_validation_set(in_list: Sequence, fold: int, index: int):
    """train_mil.py:99-104: chunks of int(len/fold) items; chunk `index` is the test set, the rest (including
    the short tail chunk the integer division leaves) the training set."""
    items = list(in_list)
    n = int(len(items) / fold)
    if n < 1:
        raise ValueError(f"{len(items)} bags cannot be split into {fold} folds")
    chunks = [items[i:i + n] for i in range(0, len(items), n)]
    test = chunks.pop(index)
    return [x for c in chunks for x in c], test


def compute_pos_weight(bags: Sequence[Tuple[int, object]]) -> float:
    """train_mil.py:106-110: negatives / positives over the bag labels (clipped to {0, 1})."""
    pos = sum(int(min(max(b[0], 0), 1)) for b in bags)
    if pos == 0:
        raise ZeroDivisionError("no positive bag in the training split (the reference divides by zero here too)")
    return (len(bags) - pos) / pos


def mil_epoch_train(milnet, store: DeviceBagStore, criterion, optimizer, order: Optional[Sequence[int]] = None) -> float:
    """train_mil.epoch_train (train_mil.py:42-59): per bag, shuffle the instances, forward, 0.5/0.5 loss, step."""
    return train_epoch(milnet, store, criterion, optimizer, dropout_patch=0.0, order=order)


def mil_epoch_test(milnet, store: DeviceBagStore, criterion, bags_per_launch: int = 16):
    """train_mil.epoch_test (train_mil.py:61-80): (mean loss, bag labels, sigmoid bag predictions)."""
    loss, labels, preds = eval_epoch(milnet, store, criterion, average=False, bags_per_launch=bags_per_launch)
    return loss, [int(l[0]) for l in labels], [p.squeeze() for p in preds]
