"""On-disk formats either side of the aggregator, pinned against the reference's own
readers: tests/golden/formats/expected.npz holds what train_tcga.get_bag_feats / generate_pt_files and
train_mil.get_data / get_bag returned for the committed input files (oracle/gen_format_golden.py)."""
import os

import numpy as np
import pytest
import torch

from dsmil_wsi_b200 import formats as F

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formats")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(FIX, "expected.npz")))


@pytest.mark.parametrize("stem", ["bag_a", "bag_b", "bag_c"])
def test_bag_csv_values_bit_identical_to_reference_reader(gold, stem):
    x = F.read_bag_csv(os.path.join(FIX, stem + ".csv"))
    assert x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]
    assert np.array_equal(x, gold[f"feats_{stem}"])


def test_bag_csv_shuffle_is_a_row_permutation(gold):
    p = os.path.join(FIX, "bag_c.csv")
    x = F.read_bag_csv(p, shuffle_rows=True, rng=np.random.default_rng(0))
    ref = gold["feats_bag_c"]
    assert x.shape == ref.shape and not np.array_equal(x, ref)
    assert np.array_equal(x[np.lexsort(x.T[::-1])], ref[np.lexsort(ref.T[::-1])])


def test_index_and_label_vectors_match_reference(gold):
    rows = F.read_dataset_index(os.path.join(FIX, "index.csv"))
    assert [r[0] for r in rows] == ["bag_a.csv", "bag_b.csv", "bag_c.csv"] and [r[1] for r in rows] == [0, 1, 2]
    for C in (1, 2, 3, 4):
        for path, label in rows:
            stem = os.path.splitext(path)[0]
            got = F.bag_label(label, C)
            assert got.dtype == np.float32 and np.array_equal(got, gold[f"label_C{C}_{stem}"]), (C, stem)
    # label >= C is the all-zero "negative" bag (train_tcga.py:31-32)
    assert not F.bag_label(2, 2).any()


def test_tcga_default_path_rewrite():
    assert F.tcga_default_feats_path("x/TCGA-05-4244-01Z-00-DX1/y") == \
        "datasets/tcga-dataset/tcga_lung_data_feats/TCGA-05-4244-01Z-00-DX1.csv"


def test_pt_cache_matches_reference(gold, tmp_path, monkeypatch):
    monkeypatch.chdir(FIX)
    out = F.generate_pt_files("index.csv", 3, out_dir=str(tmp_path / "temp_train"), shuffle_rows=False)
    assert sorted(os.path.basename(p) for p in out) == ["bag_a.pt", "bag_b.pt", "bag_c.pt"]
    for p in out:
        st = torch.load(p)
        ref = gold["pt_C3_" + os.path.splitext(os.path.basename(p))[0]]
        assert st.dtype == torch.float32 and np.array_equal(st.numpy(), ref)
        feats, label = F.split_stacked(st, 5)
        assert feats.shape == (ref.shape[0], 5) and label.shape == (1, 3)
        assert np.array_equal(label.numpy()[0], ref[0, 5:])


def test_dataset_index_writer_round_trip(tmp_path):
    root = tmp_path / "datasets" / "toy"
    for cls, names in (("b_tumor", ["s3", "s4"]), ("a_normal", ["s1"])):
        os.makedirs(root / cls)
        for n in names:
            (root / cls / (n + ".csv")).write_text("0,1\n0.5,0.25\n")
    idx = F.write_dataset_index(str(root), "toy", shuffle=True, rng=np.random.default_rng(1))
    rows = F.read_dataset_index(idx)
    # class folders sorted alphabetically give the label index (compute_feats.py:249-251)
    assert sorted((os.path.basename(p), l) for p, l in rows) == [("s1.csv", 0), ("s3.csv", 1), ("s4.csv", 1)]
    assert sorted(os.path.basename(p) for p, _ in F.read_dataset_index(str(root / "b_tumor.csv"))) == ["s3.csv", "s4.csv"]
    for p, _ in rows:
        assert F.read_bag_csv(p).tolist() == [[0.5, 0.25]]


def test_svm_reader_matches_reference_quirks(gold):
    data = F.read_mil_svm(os.path.join(FIX, "toy.svm"))
    assert np.array_equal(np.array([[d[0], d[1], d[2]] for d in data]), gold["svm_ids"])
    assert np.array_equal(np.array([len(d[3]) for d in data]), gold["svm_len"])
    vals = np.concatenate([d[3] for d in data])
    assert vals.dtype == np.float64 and np.array_equal(vals, gold["svm_vals"])
    # first line swallowed as header; trailing space gives a 5th zero slot; index before ':' ignored
    assert data[0][0] == 1 and gold["svm_len"].max() == 5 and data[1][3][1] == 0.5
    # a ragged bag fails in np.stack exactly as train_mil.py:48 would
    with pytest.raises(ValueError, match="same shape"):
        F.mil_bags(data, num_feats=4)
    for d in data:
        d[3] = d[3][:4]
    bags = F.mil_bags(data, num_feats=4)
    assert len(bags) == int(gold["svm_num_bag"])
    for b, (label, x) in enumerate(bags):
        assert x.shape == (int(gold[f"svm_bag{b}_n"]), 4) and x.dtype == np.float32
        assert label == int(np.clip(gold[f"svm_bag{b}_label"], 0, 1))


def test_svm_writer_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    bags = [(int(rng.integers(0, 2)), rng.standard_normal((int(rng.integers(1, 6)), 6)).astype(np.float32))
            for _ in range(9)]
    p = str(tmp_path / "m.svm")
    F.write_mil_svm(p, bags)
    back = F.mil_bags(F.read_mil_svm(p), num_feats=6)
    assert len(back) == len(bags)
    for (l0, x0), (l1, x1) in zip(bags, back):
        assert l0 == l1 and np.array_equal(x0, x1)
    with pytest.raises(ValueError, match="no instances"):
        d = F.read_mil_svm(p)
        F.mil_bags([r for r in d if r[1] != 2])


@pytest.mark.parametrize("N,D,C", [(0, 8, 0), (1, 1, 1), (37, 166, 1), (300, 512, 2), (5, 1024, 7)])
def test_bag_container_round_trip(tmp_path, N, D, C):
    rng = np.random.default_rng(N + D)
    x = rng.standard_normal((N, D)).astype(np.float32)
    y = rng.random(C).astype(np.float32) if C else None
    p = str(tmp_path / "bag.bin")
    F.write_bag_bin(p, x, y)
    n, d, c, off = F.read_bag_header(p)
    assert (n, d, c) == (N, D, C) and off % 64 == 0 and os.path.getsize(p) == off + 4 * N * D
    feats, label = F.read_bag_bin(p)
    assert feats.dtype == torch.float32 and feats.is_contiguous() and np.array_equal(feats.numpy(), x)
    assert np.array_equal(label.numpy(), y if C else np.zeros(0, np.float32))
    slot = torch.full((N * D + 5,), -1.0)
    feats2, _ = F.read_bag_bin(p, out=slot)
    assert (N == 0 or feats2.data_ptr() == slot.data_ptr()) and torch.equal(feats2, feats) and (slot[N * D:] == -1).all()


def test_bag_container_rejects_foreign_and_damaged_files(tmp_path):
    p = str(tmp_path / "bag.bin")
    F.write_bag_bin(p, np.ones((4, 3), np.float32), [1.0])
    raw = open(p, "rb").read()
    bad = {"magic": b"NOTABAG!" + raw[8:], "version": raw[:8] + (2).to_bytes(4, "little") + raw[12:],
           "truncated payload": raw[:-4], "truncated header": raw[:10]}
    for what, blob in bad.items():
        q = str(tmp_path / "bad.bin")
        open(q, "wb").write(blob)
        with pytest.raises(ValueError):
            F.read_bag_bin(q)
    with pytest.raises(ValueError, match="at least"):
        F.read_bag_bin(p, out=torch.empty(3))
    with pytest.raises(ValueError, match=r"\[N, D\]"):
        F.write_bag_bin(p, np.ones(4, np.float32))


def test_csv_to_container_keeps_reference_values(gold, tmp_path):
    p = str(tmp_path / "c.bin")
    assert F.csv_to_bin(os.path.join(FIX, "bag_c.csv"), p, label=[0, 0, 1]) == (23, 5)
    feats, label = F.read_bag_bin(p)
    assert np.array_equal(feats.numpy(), gold["feats_bag_c"]) and label.tolist() == [0, 0, 1]


def test_device_store_routes_agree(gold, tmp_path, monkeypatch):
    """.pt cache (reference route), index+CSV, and containers all give the store the same (feats, label) bags."""
    from dsmil_wsi_b200.feed import DeviceBagStore
    monkeypatch.chdir(FIX)
    pts = F.generate_pt_files("index.csv", 3, out_dir=str(tmp_path / "temp_train"), shuffle_rows=False)
    a, b, c = DeviceBagStore(5, device="cpu"), DeviceBagStore(5, device="cpu"), DeviceBagStore(5, device="cpu")
    a.add_files(pts)
    b.add_index("index.csv", 3)
    bins = []
    for (path, label) in F.read_dataset_index("index.csv"):
        q = str(tmp_path / (os.path.splitext(path)[0] + ".bin"))
        F.csv_to_bin(path, q, F.bag_label(label, 3))
        bins.append(q)
    c.add_bins(bins)
    assert len(a) == len(b) == len(c) == 3
    for (fa, la), (fb, lb), (fc, lc) in zip(a.bags, b.bags, c.bags):
        assert torch.equal(fa, fb) and torch.equal(fa, fc) and fa.is_contiguous() and fc.is_contiguous()
        assert la.shape == (1, 3) and torch.equal(la, lb) and torch.equal(la, lc)
    with pytest.raises(ValueError, match=r"\[N, 5\]"):
        a.add_bag(torch.zeros(3, 4), torch.zeros(3))


def test_container_writer_of_the_embedding_loop(tmp_path):
    from dsmil_wsi_b200.embed import write_bag_container, write_bag_csv
    feats = np.random.default_rng(0).random((6, 8)).astype(np.float32)
    bag_dir = os.path.join("WSI", "ds", "single", "tumor", "slide_7")
    p_bin = write_bag_container(feats, str(tmp_path), bag_dir)
    p_csv = write_bag_csv(feats, str(tmp_path), bag_dir)
    assert os.path.splitext(p_bin)[0] == os.path.splitext(p_csv)[0] and p_bin.endswith(os.path.join("tumor", "slide_7.bin"))
    exact, _ = F.read_bag_bin(p_bin)
    assert np.array_equal(exact.numpy(), feats)                        # lossless
    assert np.abs(F.read_bag_csv(p_csv) - feats).max() <= 5.1e-5       # the CSV keeps 4 decimals


# ---- native bag-CSV reader / writer (csrc_host/bagcsv.c) vs the reference's pandas route --------------------------


def _py_format(feats):
    lines = [",".join(str(i) for i in range(feats.shape[1]))]
    lines += [",".join("%.4f" % v for v in row) for row in feats]
    return "\n".join(lines) + "\n"


def test_native_writer_is_python_percent_4f_for_every_float32_it_meets():
    """'%.4f' % float(v): exact decimal of the binary value, ties to even, sign of -0.0000 kept."""
    from dsmil_wsi_b200.embed import format_bag_csv
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, size=120_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    bits = bits[np.isfinite(bits)]
    k = rng.integers(0, 200_000, size=40_000)
    vals = np.concatenate([
        bits,                                                          # any finite float32 incl. denormals, 1e38
        (np.abs(rng.standard_normal(60_000)) * 2).astype(np.float32),  # the range features live in
        ((2 * k + 1) / 20000.0).astype(np.float32),                    # neighbours of the x.xxxx5 ties
        np.array([0.0, -0.0, 0.00005, 0.00015, -0.00005, 0.99995, 12345.678, 2.0 ** 39, -2.0 ** 39, 2.0 ** 40,
                  3.4e38, 1e-45, np.inf, -np.inf], np.float32)])
    vals = vals[: vals.size // 7 * 7].reshape(-1, 7)
    assert format_bag_csv(vals) == _py_format(vals)


def test_native_writer_matches_pandas_text_and_nan_convention(tmp_path):
    import pandas as pd
    from dsmil_wsi_b200.embed import format_bag_csv, write_bag_csv
    rng = np.random.default_rng(1)
    feats = (rng.standard_normal((333, 17)) * np.array([1e-5, 1, 10, 1000] * 4 + [1e6])).astype(np.float32)
    feats[3, 4] = np.nan                                               # pandas writes an empty field
    feats[0, 0], feats[1, 1] = 0.0, -0.00004
    buf = __import__("io").StringIO()
    pd.DataFrame(feats).to_csv(buf, index=False, float_format="%.4f")
    assert format_bag_csv(feats) == buf.getvalue()
    p = write_bag_csv(feats, str(tmp_path), os.path.join("WSI", "ds", "single", "c0", "s1"))
    assert open(p).read() == buf.getvalue()
    for N in (0, 1):                                                    # degenerate bags still get the header
        buf = __import__("io").StringIO()
        pd.DataFrame(np.zeros((N, 5), np.float32), columns=range(5)).to_csv(buf, index=False, float_format="%.4f")
        assert format_bag_csv(np.zeros((N, 5), np.float32)) == buf.getvalue()


@pytest.mark.parametrize("N,D", [(1, 1), (50, 512), (1000, 7)])
def test_native_reader_is_bit_identical_to_the_pandas_route(tmp_path, N, D):
    from dsmil_wsi_b200.embed import format_bag_csv
    rng = np.random.default_rng(N + D)
    feats = (np.abs(rng.standard_normal((N, D))) * rng.choice([1e-3, 1.0, 50.0, 2e4], size=(N, D))).astype(np.float32)
    p = str(tmp_path / "bag.csv")
    open(p, "w").write(format_bag_csv(feats))
    a, b = F.read_bag_csv(p), F.read_bag_csv(p, engine="pandas")
    assert a.dtype == np.float32 and a.shape == (N, D) and np.array_equal(a, b)
    assert np.abs(a - feats).max() <= 5.1e-5 * max(1.0, np.abs(feats).max())
    out = np.full(N * D + 3, -1, np.float32)
    c = F.read_bag_csv(p, out=out)
    assert np.shares_memory(c, out) and np.array_equal(c, a) and (out[N * D:] == -1).all()


def test_native_reader_general_fields_and_errors(tmp_path):
    """Hand-written CSVs: exponents, signs, long mantissas, CRLF, blank lines, empty fields -- against pandas."""
    text = "0,1,2\r\n1e-3,-2.5E+2,+7\r\n\r\n0.1234567890123456789,123456789012345678901,.5\r\n,3.,-0\r\n"
    p = str(tmp_path / "odd.csv")
    open(p, "w", newline="").write(text)
    a, b = F.read_bag_csv(p), F.read_bag_csv(p, engine="pandas")
    assert a.shape == b.shape == (3, 3)
    assert np.array_equal(a, b, equal_nan=True) and np.isnan(a[2, 0]) and np.signbit(a[2, 2])
    for bad, needle in (("0,1\n1,2,3\n", "ragged"), ("0,1\n1\n", "ragged"), ("0,1\n1,abc\n", "not a number"),
                        ("", "not a bag CSV")):
        q = str(tmp_path / "bad.csv")
        open(q, "w").write(bad)
        with pytest.raises(ValueError, match=needle):
            F.read_bag_csv(q)
    with pytest.raises(ValueError, match="engine"):
        F.read_bag_csv(p, engine="polars")


def test_native_reader_fast_path_is_correctly_rounded(tmp_path):
    """Fields of up to 18 significant digits: float32(correctly rounded double) == np.float32(float(text))."""
    rng = np.random.default_rng(5)
    fields = []
    for _ in range(20_000):
        nd = int(rng.integers(1, 19))
        digits = "".join(str(d) for d in rng.integers(0, 10, size=nd))
        cut = int(rng.integers(0, nd + 1))
        text = (digits[:cut] or "0") + ("." + digits[cut:] if cut < nd else "")
        fields.append(("-" if rng.random() < 0.3 else "") + text)
    D = 8
    rows = [fields[i:i + D] for i in range(0, len(fields), D)]
    p = str(tmp_path / "f.csv")
    open(p, "w").write(",".join(str(i) for i in range(D)) + "\n" + "\n".join(",".join(r) for r in rows) + "\n")
    got = F.read_bag_csv(p)
    want = np.array([[np.float32(float(t)) for t in r] for r in rows], dtype=np.float32)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(np.signbit(got), np.signbit(want))
