"""Pins oracle/dsmil_oracle.py against fixtures produced by the unmodified reference
(oracle/gen_golden.py; reference dsmil.py:10-12,46-62,70-74 + train_tcga.py:67-72 loss)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_to_max
from oracle import dsmil_oracle as orc

# The fixtures are fp32 outputs of torch-CPU; vs our fp64 restatement the distance is the
# reference's own fp32 noise (classes 5e-7, logits up to 1.5e-5, A 3e-6, B 1e-6).
TOL = dict(classes=2e-6, A=1e-5, B=4e-6, pred=3e-5)


@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference(name):
    g, p, X = load_golden(name)
    for dt in (np.float64, np.float32):
        out = orc.forward(X, p, dtype=dt)
        assert np.array_equal(out.idx, g["idx"]), (name, out.idx, g["idx"])  # arg-max indices bit-exact
        assert out.classes.shape == g["classes"].shape and out.A.shape == g["A"].shape
        assert out.B.shape == g["B"].shape and out.prediction_bag.shape == g["pred"].shape
        assert rel_to_max(out.classes, g["classes"]) < TOL["classes"]
        assert rel_to_max(out.A, g["A"]) < TOL["A"] * (4 if dt is np.float32 else 1)
        assert rel_to_max(out.B, g["B"]) < TOL["B"] * (4 if dt is np.float32 else 1)
        # bag logits: |d| <= tol * max(|logit|, scale of the GEMV terms) -- cancellation in the fcc GEMV
        terms = np.abs(p.Wf.reshape(p.C, -1)).astype(np.float64) @ np.abs(g["B"].reshape(-1).astype(np.float64))
        err = np.abs(out.prediction_bag.reshape(-1).astype(np.float64) - g["pred"].reshape(-1))
        assert np.all(err <= TOL["pred"] * np.maximum(np.abs(g["pred"].reshape(-1)), terms)), (name, err)
        assert np.allclose(out.A.sum(0), 1.0, atol=1e-5)


@pytest.mark.parametrize("name", golden_names())
def test_backward_matches_autograd_through_reference(name):
    g, p, X = load_golden(name)
    out = orc.forward(X, p, dtype=np.float64)
    loss, d_cls, d_pred = orc.caller_loss_grads(out, g["y"])
    assert abs(loss - float(g["loss"])) < 2e-6 * max(1.0, abs(loss))
    grads = orc.backward(X, p, out, d_cls, d_pred, need_dX="g_X" in g)
    for k, v in grads.items():
        ref = g["g_" + k]
        assert v.shape == ref.shape, (k, v.shape, ref.shape)
        r = rel_to_max(v, ref)
        assert r < 3e-4, (name, k, r)   # fp32 autograd noise on 1e-5-magnitude q.* grads


@pytest.mark.parametrize("name", ["shipped_tcga", "rand_d512_c1", "musk_d166_n7", "pv_d96_c2", "lin_d512_c3"])
@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_sharded_algebra_reproduces_single_device(name, G):
    g, p, X = load_golden(name)
    one = orc.forward(X, p, dtype=np.float64)
    sh = orc.forward_sharded(X, p, G, dtype=np.float64)
    assert np.array_equal(one.idx, sh.idx)
    for f in ("classes", "A", "B", "prediction_bag"):
        assert rel_to_max(getattr(sh, f), getattr(one, f)) < 1e-12, f


def test_sharded_more_ranks_than_rows():
    g, p, X = load_golden("musk_d166_n2")
    one = orc.forward(X, p)
    sh = orc.forward_sharded(X, p, 8)
    assert np.array_equal(one.idx, sh.idx) and rel_to_max(sh.A, one.A) < 1e-12


def test_torch_port_matches_golden():
    g, p, X = load_golden("shipped_tcga")
    import torch
    port = orc.TorchPort(p, threads=2)
    c, pred, A, B = port.forward(torch.from_numpy(X))
    assert rel_to_max(c.numpy(), g["classes"]) < 2e-6
    assert rel_to_max(A.numpy(), g["A"]) < 1e-5
    assert rel_to_max(B.numpy(), g["B"]) < 4e-6
    assert pred.shape == (1, 2) and B.shape == (1, 2, 512)


def test_tie_and_nan_semantics():
    c = np.array([[1.0, 5.0], [3.0, 5.0], [3.0, 2.0]])
    assert list(orc.critical_instances(c)) == [1, 0]          # lowest index wins ties
    c[2, 0] = np.nan
    assert orc.critical_instances(c)[0] == 2                   # NaN ranks first (torch.sort descending)


def test_split_precision_emulation_is_at_fp32_floor():
    """3xBF16 (hi*hi + lo*hi + hi*lo) is what the tensor-core kernel computes; its error on the
    Q-MLP pre-activations must sit at the fp32 noise floor."""
    g, p, X = load_golden("shipped_tcga")
    truth = X.astype(np.float64) @ p.W1.astype(np.float64).T
    e3 = rel_to_max(orc.matmul_3xbf16(X, p.W1), truth)
    e1 = rel_to_max(orc.bf16_round(X).astype(np.float64) @ orc.bf16_round(p.W1).astype(np.float64).T, truth)
    f32 = rel_to_max((X @ p.W1.T).astype(np.float64), truth)
    assert e3 < 6e-6 and e1 > 50 * e3, (e3, e1, f32)


@pytest.mark.parametrize("name", ["shipped_tcga", "shipped_c16", "rand_d512_c2", "rand_d512_c1"])
def test_apriori_softmax_bound_on_tanh_path(name):
    """with a tanh-bounded Q, |L| <= ||q_max||_1 / sqrt(128f), so exp(L - bound) needs no running
    max; A and B from plain sums stay at the oracle's fp32 noise floor."""
    from conftest import load_golden, rel_to_max
    g, p, X = load_golden(name)
    X = X.astype(np.float32)
    Q = orc.q_mlp(X, p)[0].astype(np.float32)
    idx = orc.critical_instances(X @ p.Wi.T + p.bi)
    qmax = Q[idx]
    s = np.float32(np.sqrt(np.float32(128)))
    L = (Q @ qmax.T) / s
    bound = (np.abs(qmax).sum(1) / s).astype(np.float32)
    assert (np.abs(L) <= bound[None, :] * (1 + 1e-6)).all() and bound.max() <= 11.3138
    e = np.exp((L - bound[None, :]).astype(np.float32))
    assert e.min() > 1e-10                                   # far from denormals: e^-22.7 = 1.4e-10 is the floor
    S = e.sum(0, dtype=np.float64)
    A = (e / S).astype(np.float32)
    B = (e.astype(np.float64).T @ X.astype(np.float64)) / S[:, None]
    assert rel_to_max(A, g["A"]) < 5e-6 and rel_to_max(B, g["B"].reshape(B.shape)) < 5e-6


@pytest.mark.parametrize("name", ["shipped_tcga", "rand_d512_c1", "musk_d166_n7", "lin_d512_c3"])
@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_sharded_backward_algebra_reproduces_single_device(name, G):
    """SURVEY A.2 "Sharded": three reductions (t, dq_max, parameter grads) give the single-device gradients."""
    from conftest import load_golden, rel_to_max
    g, p, X = load_golden(name)
    out = orc.forward(X, p)
    y = np.asarray(g["y"], np.float64).reshape(-1)
    _, d_cls, d_pred = orc.caller_loss_grads(out, y)
    one = orc.backward(X, p, out, d_cls, d_pred)
    many = orc.backward_sharded(X, p, out, d_cls, d_pred, G)
    assert set(many) == {k for k in one if k != "X"}
    for k in many:
        assert rel_to_max(many[k], one[k]) < 1e-12, k
