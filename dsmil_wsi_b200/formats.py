"""On-disk formats either side of the aggregator.

Readers/writers for the four formats the reference's callers exchange, with the reference's exact value
semantics, plus a binary bag container that skips the `%.4f` text round trip when producer and consumer are
both ours:

    bag feature CSV      compute_feats.py:80-82,123-125  ->  train_tcga.py:24-26
    dataset index CSV    compute_feats.py:249-260        ->  train_tcga.py:245-250,19-34
    training cache .pt   train_tcga.py:36-51             ->  train_tcga.py:62-64,93-95
    classic-MIL svm text (external)                       ->  train_mil.py:17-40,144-149

Host-side only (numpy / pandas / torch CPU): nothing here touches the device; `feed.DeviceBagStore` moves the
results into HBM.  Pinned against the reference's own readers by tests/test_formats.py
(fixtures: oracle/gen_format_golden.py).
"""
from __future__ import annotations

import glob
import os
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

# ---- bag feature CSV ---------------------------------------------------------------------------


def read_bag_csv(path: str, shuffle_rows: bool = False, rng: Optional[np.random.Generator] = None,
                 engine: str = "native", out: Optional[np.ndarray] = None) -> np.ndarray:
    """[N, D] float32 features of one bag (train_tcga.py:24-26,46).

    The reference parses with `pd.read_csv` (float64) and rounds to fp32 in `torch.tensor(..., float32)`.
    engine="native" (default) is the C parser of csrc_host/bagcsv.c -- the same values bit for bit for the
    `%.4f` files the embedding loop writes (and any field of up to 18 significant digits), about ten times
    faster; engine="pandas" is the reference's own route, kept for cross-checks.  The first line is the
    `0..D-1` header.  The reference shuffles rows on every read (`sklearn.utils.shuffle`, unseeded); that is
    opt-in here because the training feed permutes on the device per epoch anyway (feed.dropout_patches).
    `out`: optional preallocated C-contiguous float32 buffer of at least N*D elements (e.g. pinned memory)."""
    if engine == "pandas":
        import pandas as pd
        feats = pd.read_csv(path).to_numpy()
        if feats.ndim != 2:
            raise ValueError(f"{path}: expected a 2-D table, got shape {feats.shape}")
        feats = np.ascontiguousarray(feats, dtype=np.float32)
    elif engine == "native":
        feats = _read_bag_csv_native(path, out)
    else:
        raise ValueError(f"engine must be 'native' or 'pandas', got {engine!r}")
    if shuffle_rows:
        rng = rng or np.random.default_rng()
        feats = feats[rng.permutation(feats.shape[0])]
    return feats


def _read_bag_csv_native(path: str, out: Optional[np.ndarray]) -> np.ndarray:
    import ctypes as C
    from . import _hostlib
    lib = _hostlib.load()
    raw = np.fromfile(path, dtype=np.uint8)
    N, D = C.c_int64(0), C.c_int32(0)
    rc = lib.dsmil_csv_shape(raw.ctypes.data, raw.size, C.byref(N), C.byref(D))
    if rc:
        raise ValueError(f"{path}: not a bag CSV ({_hostlib.ERRORS.get(rc, rc)})")
    n, d = int(N.value), int(D.value)
    if out is not None:
        if out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or out.size < n * d:
            raise ValueError("out must be a C-contiguous float32 array with at least N*D elements")
        feats = out.reshape(-1)[: n * d].reshape(n, d)
    else:
        feats = np.empty((n, d), dtype=np.float32)
    bad = C.c_int64(0)
    rc = lib.dsmil_csv_parse_bag(raw.ctypes.data, raw.size, feats.ctypes.data, n, d, C.byref(bad))
    if rc:
        raise ValueError(f"{path}: data line {bad.value}: {_hostlib.ERRORS.get(rc, rc)}")
    return feats


def bag_label(label, num_classes: int) -> np.ndarray:
    """Label column of the dataset index -> target vector (train_tcga.py:27-32).

    C == 1: the value itself.  C > 1: one-hot at int(label); a label >= C gives the all-zero "negative" bag.
    """
    out = np.zeros(num_classes, dtype=np.float32)
    if num_classes == 1:
        out[0] = label
    elif int(label) <= num_classes - 1:
        out[int(label)] = 1
    return out


# ---- dataset index CSV -------------------------------------------------------------------------


def read_dataset_index(path: str) -> List[Tuple[str, float]]:
    """Rows of `datasets/<name>/<name>.csv`: (bag csv path, label) (train_tcga.py:245-250 -> :19-34).

    Column 0 is the path (header `0`), column 1 the class-folder index (header `label`)."""
    import pandas as pd
    df = pd.read_csv(path)
    if df.shape[1] < 2:
        raise ValueError(f"{path}: a dataset index has two columns (path, label), found {df.shape[1]}")
    return [(str(r.iloc[0]), r.iloc[1].item() if hasattr(r.iloc[1], "item") else r.iloc[1]) for _, r in df.iterrows()]


def tcga_default_feats_path(index_path_entry: str) -> str:
    """The path rewrite of the `TCGA-lung-default` dataset (train_tcga.py:20-21)."""
    return "datasets/tcga-dataset/tcga_lung_data_feats/" + index_path_entry.split("/")[1] + ".csv"


def write_dataset_index(dataset_dir: str, name: str, shuffle: bool = True,
                        rng: Optional[np.random.Generator] = None) -> str:
    """compute_feats.py:249-260: one `<class>.csv` per class folder (sorted => label index) and the
    concatenated `<name>.csv`, rows shuffled.  Returns the path of the latter."""
    classes = sorted(glob.glob(os.path.join(dataset_dir, "*" + os.path.sep)))
    rows: List[Tuple[str, int]] = []
    for i, item in enumerate(classes):
        csvs = glob.glob(os.path.join(item, "*.csv"))
        cls = os.path.basename(os.path.normpath(item))
        with open(os.path.join(dataset_dir, cls + ".csv"), "w") as f:
            f.write("0,label\n" + "".join(f"{p},{i}\n" for p in csvs))
        rows += [(p, i) for p in csvs]
    if shuffle:
        rng = rng or np.random.default_rng()
        rows = [rows[j] for j in rng.permutation(len(rows))]
    out = os.path.join(dataset_dir, name + ".csv")
    with open(out, "w") as f:
        f.write("0,label\n" + "".join(f"{p},{i}\n" for p, i in rows))
    return out


# ---- training cache (.pt) ----------------------------------------------------------------------


def stack_bag(feats: np.ndarray, label: np.ndarray) -> torch.Tensor:
    """[N, D + C] fp32 = features || label repeated per row (train_tcga.py:45-49)."""
    f = torch.as_tensor(np.asarray(feats), dtype=torch.float32)
    y = torch.as_tensor(np.asarray(label), dtype=torch.float32).view(1, -1)
    return torch.cat((f, y.repeat(f.size(0), 1)), dim=1)


def generate_pt_files(index_path: str, num_classes: int, out_dir: str = "temp_train",
                      tcga_default: bool = False, shuffle_rows: bool = True) -> List[str]:
    """train_tcga.py:36-51 without the `rmtree` of the cwd-relative directory being implicit: `out_dir` is
    created if absent and files are overwritten.  File name = bag csv stem + '.pt'."""
    os.makedirs(out_dir, exist_ok=True)
    out = []
    for entry, label in read_dataset_index(index_path):
        csv = tcga_default_feats_path(entry) if tcga_default else entry
        st = stack_bag(read_bag_csv(csv, shuffle_rows=shuffle_rows), bag_label(label, num_classes))
        p = os.path.join(out_dir, os.path.splitext(csv)[0].split(os.sep)[-1] + ".pt")
        torch.save(st, p)
        out.append(p)
    return out


def split_stacked(stacked: torch.Tensor, feats_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(bag_feats [N, D], bag_label [1, C]) of a cache tensor (train_tcga.py:62-64)."""
    return stacked[:, :feats_size], stacked[0, feats_size:].unsqueeze(0)


# ---- classic-MIL svm text ----------------------------------------------------------------------


def read_mil_svm(path: str) -> List[list]:
    """`[inst_id, bag_id, label, feature_vector(float64)]` per line, exactly as train_mil.py:17-35 yields it.

    Quirks kept on purpose (they define which instances the reference trains on):
      * the FIRST line is consumed as a CSV header by `pd.read_csv` and never becomes an instance;
      * lines are split on single spaces; token j (0-based, after the id token) fills slot j of a vector as
        long as the token count -- the feature INDEX before ':' is ignored, a token without exactly one ':'
        (e.g. the empty token of a trailing space) leaves a 0;
      * only the text up to the first comma would survive `df[df.columns[0]]`; svm files have no commas.
    """
    out = []
    with open(path, "r") as f:
        lines = f.read().splitlines()
    for line in lines[1:]:
        if line == "":
            continue                       # pandas skips blank lines
        line = line.split(",")[0]
        toks = line.split(" ")
        ids = toks[0].split(":")
        vec = np.zeros(len(toks) - 1)
        for j, t in enumerate(toks[1:]):
            kv = t.split(":")
            if len(kv) == 2:
                vec[j] = float(kv[1])
        out.append([int(ids[0]), int(ids[1]), int(ids[2]), vec])
    return out


def mil_bags(data: Sequence[list], num_feats: Optional[int] = None) -> List[Tuple[int, np.ndarray]]:
    """Group instances into bags (train_mil.py:37-40,144-149): `num_bag = last line's bag id + 1`, the bag
    label is its first instance's label clipped to {0,1} (`np.clip(label, 0, 1)` at :49,:69), features are
    stacked to [n_i, D] and cut to `num_feats` columns (:48)."""
    if not data:
        return []
    num_bag = data[-1][1] + 1
    bags = []
    for b in range(num_bag):
        rows = [d for d in data if d[1] == b]
        if not rows:
            raise ValueError(f"bag id {b} has no instances (ids must be contiguous from 0, train_mil.py:145-147)")
        x = np.stack([r[3] for r in rows])
        if num_feats is not None:
            x = x[:, :num_feats]
        bags.append((int(np.clip(rows[0][2], 0, 1)), x.astype(np.float32)))
    return bags


def write_mil_svm(path: str, bags: Sequence[Tuple[int, np.ndarray]], header: str = "# synthetic") -> None:
    """Writer for synthetic fixtures in the format above (labels written as -1/+1 like the UCI files)."""
    inst = 0
    with open(path, "w") as f:
        f.write(header + "\n")
        for b, (label, x) in enumerate(bags):
            for row in np.asarray(x):
                feats = " ".join(f"{j}:{repr(float(v))}" for j, v in enumerate(row))
                f.write(f"{inst}:{b}:{1 if label > 0 else -1} {feats}\n")
                inst += 1


# ---- binary bag container ----------------------------------------------------------------------
#
#   offset  size  field
#   0       8     magic  b"DSMILBAG"
#   8       4     version (1)                      little-endian throughout
#   12      4     D  (features per instance)
#   16      8     N  (instances)
#   24      4     C  (label length, may be 0)
#   28      4     data offset in bytes (multiple of 64)
#   32      4*C   label, fp32
#   off     4*N*D features, fp32 row-major -- the layout `dsmil_forward` reads, so the payload can be read
#                 straight into a pinned staging buffer and copied to HBM with no parse or transpose.

BAG_MAGIC = b"DSMILBAG"
BAG_VERSION = 1
_HDR = struct.Struct("<8sIIQII")


def write_bag_bin(path: str, feats, label=None) -> None:
    x = np.ascontiguousarray(np.asarray(feats), dtype="<f4")
    if x.ndim != 2:
        raise ValueError(f"feats must be [N, D], got shape {x.shape}")
    y = np.zeros(0, "<f4") if label is None else np.ascontiguousarray(np.asarray(label).reshape(-1), dtype="<f4")
    off = -(-(_HDR.size + 4 * y.size) // 64) * 64
    with open(path, "wb") as f:
        f.write(_HDR.pack(BAG_MAGIC, BAG_VERSION, x.shape[1], x.shape[0], y.size, off))
        f.write(y.tobytes())
        f.write(b"\0" * (off - _HDR.size - 4 * y.size))
        f.write(x.tobytes())


def read_bag_header(path: str) -> Tuple[int, int, int, int]:
    """(N, D, C, data offset); raises ValueError on anything that is not a version-1 container."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        raw = f.read(_HDR.size)
    if len(raw) < _HDR.size:
        raise ValueError(f"{path}: truncated header ({len(raw)} bytes)")
    magic, ver, D, N, C, off = _HDR.unpack(raw)
    if magic != BAG_MAGIC:
        raise ValueError(f"{path}: not a DSMIL bag container (magic {magic!r})")
    if ver != BAG_VERSION:
        raise ValueError(f"{path}: container version {ver}, this reader understands {BAG_VERSION}")
    if off % 64 or off < _HDR.size + 4 * C:
        raise ValueError(f"{path}: bad data offset {off}")
    if size != off + 4 * N * D:
        raise ValueError(f"{path}: size {size} B does not match header (N={N}, D={D}: expected {off + 4 * N * D} B)")
    return N, D, C, off


def read_bag_bin(path: str, pin_memory: bool = False, out: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(feats [N, D] fp32, label [C] fp32) as CPU tensors.  `pin_memory=True` reads the payload directly into
    page-locked memory (needs a CUDA runtime) so the later H2D copy is asynchronous; `out` reuses a caller
    buffer of at least N*D elements (e.g. a pinned staging slot)."""
    N, D, C, off = read_bag_header(path)
    if out is not None:
        if out.dtype != torch.float32 or out.device.type != "cpu" or not out.is_contiguous() or out.numel() < N * D:
            raise ValueError("out must be a contiguous CPU float32 tensor with at least N*D elements")
        feats = out.view(-1)[: N * D].view(N, D)
    else:
        feats = torch.empty((N, D), dtype=torch.float32, pin_memory=pin_memory)
    with open(path, "rb") as f:
        f.seek(_HDR.size)
        label = torch.from_numpy(np.frombuffer(f.read(4 * C), dtype="<f4").astype(np.float32))
        f.seek(off)
        if N * D:
            got = f.readinto(memoryview(feats.numpy()).cast("B"))
            if got != 4 * N * D:
                raise ValueError(f"{path}: short read ({got} of {4 * N * D} bytes)")
    return feats, label


def csv_to_bin(csv_path: str, bin_path: str, label=None) -> Tuple[int, int]:
    """Convert one bag feature CSV to the container (values identical to `read_bag_csv`).  Returns (N, D)."""
    x = read_bag_csv(csv_path)
    write_bag_bin(bin_path, x, label)
    return x.shape
