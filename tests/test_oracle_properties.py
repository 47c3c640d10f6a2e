"""Size-independent properties of the oracle (the checker has to be right before it checks anything):
finite differences pin `backward` independently of autograd, and the algebraic properties the GPU suite relies
on at full size (permutation equivariance, softmax simplex, convex-hull bound on B, shard-count invariance) hold
for the restatement itself on random small problems."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import dsmil_oracle as orc

SETTINGS = dict(max_examples=20, deadline=None)


def problem(seed, N, D, C, nonlinear=True, passing_v=False):
    p = orc.random_params(D, C, seed, nonlinear=nonlinear, passing_v=passing_v, scale=1.5)
    X = orc.synthetic_bag(N, D, seed + 1, "normal")
    return p, X


def gap(c):
    s = np.sort(c, 0)[::-1]
    return (s[0] - s[1]).min() if c.shape[0] > 1 else np.inf


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), N=st.integers(1, 40), D=st.sampled_from([3, 16, 33]), C=st.integers(1, 4),
       nonlinear=st.booleans(), passing_v=st.booleans())
def test_outputs_live_where_the_algebra_says(seed, N, D, C, nonlinear, passing_v):
    p, X = problem(seed, N, D, C, nonlinear, passing_v)
    out = orc.forward(X, p)
    assert out.classes.shape == (N, C) and out.prediction_bag.shape == (1, C)
    assert out.A.shape == (N, C) and out.B.shape == (1, C, D) and out.idx.shape == (C,)
    assert (out.A >= 0).all() and np.allclose(out.A.sum(0), 1.0, atol=1e-12)       # softmax over the instances
    assert np.array_equal(out.idx, out.classes.argmax(0))                           # dsmil.py:52, row 0 of the sort
    V = out.V                                                                       # B = A^T V: a convex combination
    assert (out.B[0] <= V.max(0) + 1e-12).all() and (out.B[0] >= V.min(0) - 1e-12).all()
    if nonlinear:
        assert np.abs(out.Q).max() <= 1.0 and np.abs(out.logits).max() <= 128 / orc.SCALE_F32 + 1e-9


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), N=st.integers(2, 40), C=st.integers(1, 3))
def test_row_permutation_equivariance(seed, N, C):
    p, X = problem(seed, N, 16, C)
    out = orc.forward(X, p)
    if gap(out.classes) < 1e-9:
        return                                   # a tie: which row wins is the documented lowest-index rule
    perm = np.random.default_rng(seed).permutation(N)
    outp = orc.forward(X[perm], p)
    assert np.array_equal(perm[outp.idx], out.idx)
    assert np.allclose(outp.classes, out.classes[perm], atol=1e-13) and np.allclose(outp.A, out.A[perm], atol=1e-13)
    assert np.allclose(outp.B, out.B, atol=1e-12) and np.allclose(outp.prediction_bag, out.prediction_bag, atol=1e-12)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), N=st.integers(1, 50), G=st.integers(1, 9), C=st.integers(1, 3),
       nonlinear=st.booleans())
def test_shard_count_invariance(seed, N, G, C, nonlinear):
    p, X = problem(seed, N, 16, C, nonlinear)
    one, many = orc.forward(X, p), orc.forward_sharded(X, p, G)
    if gap(one.classes) < 1e-9:
        return
    assert np.array_equal(one.idx, many.idx)
    for a, b in ((one.classes, many.classes), (one.A, many.A), (one.B, many.B), (one.prediction_bag, many.prediction_bag)):
        assert np.allclose(a, b, rtol=0, atol=1e-12)
    bounds = orc.shard_bounds(N, G)
    assert bounds[0][0] == 0 and bounds[-1][1] == N and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    assert max(hi - lo for lo, hi in bounds) - min(hi - lo for lo, hi in bounds) <= 1


def caller_loss(X, p, y, pos_weight=None):
    out = orc.forward(X, p)
    return orc.caller_loss_grads(out, y, pos_weight)[0], out


@pytest.mark.parametrize("nonlinear,passing_v,C", [(True, False, 2), (True, True, 1), (False, False, 3), (False, True, 2)])
def test_backward_against_central_differences(nonlinear, passing_v, C):
    """Every parameter tensor and dX: d loss / d theta from `backward` vs (L(theta+h) - L(theta-h)) / 2h in fp64.
    The arg-max is piecewise constant, so the step is kept far below the top-1/top-2 score gap."""
    N, D = 9, 7
    p, X = problem(3 + C, N, D, C, nonlinear, passing_v)
    y = (np.arange(C) % 2).astype(np.float64)
    loss, out = caller_loss(X, p, y, pos_weight=1.7)
    assert gap(out.classes) > 1e-3
    _, d_cls, d_pred = orc.caller_loss_grads(out, y, 1.7)
    g = orc.backward(X, p, out, d_cls, d_pred, need_dX=True)
    h = 1e-6
    names = ["Wi", "bi", "W1", "b1", "Wf", "bf"] + (["W2", "b2"] if nonlinear else []) + (["Wv", "bv"] if passing_v else [])
    rng = np.random.default_rng(0)
    for name in names + ["X"]:
        base = X if name == "X" else getattr(p, name)
        flat_idx = rng.choice(base.size, size=min(base.size, 12), replace=False)
        for k in flat_idx:
            def at(delta):
                arr = np.array(base, dtype=np.float64, copy=True)
                arr.reshape(-1)[k] += delta
                if name == "X":
                    return caller_loss(arr, p, y, 1.7)[0]
                q = p.astype(np.float64)
                setattr(q, name, arr)
                return caller_loss(X, q, y, 1.7)[0]
            fd = (at(h) - at(-h)) / (2 * h)
            an = np.asarray(g[name], np.float64).reshape(-1)[k]
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 2e-8, (name, int(k), fd, an)
