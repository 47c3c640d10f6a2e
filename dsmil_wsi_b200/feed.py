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

    def add_index(self, index_csv: str, num_classes: int, tcga_default: bool = False) -> None:
        """The reference's route (train_tcga.py:245-250 + :36-51) without the temp_train/*.pt detour."""
        from . import formats
        for entry, label in formats.read_dataset_index(index_csv):
            csv = formats.tcga_default_feats_path(entry) if tcga_default else entry
            self.add_bag(torch.from_numpy(formats.read_bag_csv(csv)),
                         torch.from_numpy(formats.bag_label(label, num_classes)))

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
