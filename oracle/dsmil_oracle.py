"""CPU oracle for the DSMIL aggregator hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``dsmil_wsi_b200/`` imports it, and the
product path raises when the CUDA library is missing instead of landing here.

Parity status: **pinned**.  The reference repo has no tests of its own
(SURVEY.md §4), so the pin is made from the reference itself: ``gen_golden.py``
imports the unmodified ``/root/reference/dsmil.py`` in the authoring container,
runs it on the shipped ``example_aggregator_weights/*.pth`` and on seeded random
weights, and stores inputs' seeds + outputs under ``tests/golden/``.
``tests/test_oracle.py`` checks every function here against those fixtures.

What is restated (reference file:line):
  * ``FCLayer.forward``                dsmil.py:10-12   -> instance scores
  * ``BClassifier.forward``            dsmil.py:46-62   -> critical instance, Q, attention,
                                                            softmax over instances, bag vector,
                                                            Conv1d bag classifier (== GEMV)
  * ``MILNet.forward``                 dsmil.py:70-74   -> return order
  * loss of the callers                train_tcga.py:67-72, train_mil.py:50-56 (for backward)
  * row-sharded combine                SURVEY.md Appendix A.3 (our own multi-GPU algebra)

Two flavours:
  * ``forward(..., dtype=np.float64)`` -- "truth": the same algebra in fp64 (the scale
    constant stays the fp32 value sqrt(128f), dsmil.py:56 builds it as a float32 tensor).
  * ``forward(..., dtype=np.float32)`` -- fp32 numpy restatement.
  * ``TorchPort`` -- the same op sequence on torch-CPU fp32 with all host threads; this is
    what ``bench.py`` times as ``cpu_baseline.kind == "port"``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

Q_DIM = 128  # hard-coded in dsmil.py:31,33
# dsmil.py:56 -- torch.sqrt(torch.tensor(128, dtype=float32)); the fp32 value, kept even in fp64 runs.
SCALE_F32 = float(np.sqrt(np.float32(Q_DIM)))


# --------------------------------------------------------------------------- parameters
@dataclass
class Params:
    """Plain-array view of a MILNet(FCLayer, BClassifier) state_dict."""

    Wi: np.ndarray  # [C, D]       i_classifier.fc.0.weight   (dsmil.py:9)
    bi: np.ndarray  # [C]
    Wf: np.ndarray  # [C, C, D]    b_classifier.fcc.weight    (dsmil.py:44)
    bf: np.ndarray  # [C]
    W1: np.ndarray  # [128, D]     q.0.weight (nonlinear) or q.weight (linear)  (dsmil.py:31,33)
    b1: np.ndarray  # [128]
    W2: Optional[np.ndarray] = None  # [128,128] q.2.weight (nonlinear only)
    b2: Optional[np.ndarray] = None
    Wv: Optional[np.ndarray] = None  # [D, D]  v.1.weight (passing_v only)  (dsmil.py:35-39)
    bv: Optional[np.ndarray] = None

    @property
    def nonlinear(self) -> bool:
        return self.W2 is not None

    @property
    def passing_v(self) -> bool:
        return self.Wv is not None

    @property
    def C(self) -> int:
        return int(self.Wi.shape[0])

    @property
    def D(self) -> int:
        return int(self.Wi.shape[1])

    def astype(self, dt) -> "Params":
        cv = lambda a: None if a is None else np.asarray(a, dtype=dt)
        return Params(cv(self.Wi), cv(self.bi), cv(self.Wf), cv(self.bf), cv(self.W1), cv(self.b1),
                      cv(self.W2), cv(self.b2), cv(self.Wv), cv(self.bv))


def params_from_state_dict(sd: Dict[str, "np.ndarray"]) -> Params:
    """Accepts the key layout the reference modules produce."""
    g = lambda k: None if k not in sd else np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k])
    wi = g("i_classifier.fc.0.weight")
    bi = g("i_classifier.fc.0.bias")
    if wi is None:  # IClassifier layout (dsmil.py:19) -- attention_map.py:163-164 re-keys to this
        wi, bi = g("i_classifier.fc.weight"), g("i_classifier.fc.bias")
    if "b_classifier.q.0.weight" in sd:
        W1, b1 = g("b_classifier.q.0.weight"), g("b_classifier.q.0.bias")
        W2, b2 = g("b_classifier.q.2.weight"), g("b_classifier.q.2.bias")
    else:
        W1, b1 = g("b_classifier.q.weight"), g("b_classifier.q.bias")
        W2 = b2 = None
    return Params(wi, bi, g("b_classifier.fcc.weight"), g("b_classifier.fcc.bias"), W1, b1, W2, b2,
                  g("b_classifier.v.1.weight"), g("b_classifier.v.1.bias"))


def random_params(D: int, C: int, seed: int, nonlinear: bool = True, passing_v: bool = False,
                  scale: float = 1.0) -> Params:
    """Seeded weights with nn.Linear-like magnitudes (uniform +-1/sqrt(fan_in))."""
    rng = np.random.default_rng(seed)
    u = lambda shape, fan: (rng.uniform(-1, 1, size=shape) * scale / math.sqrt(fan)).astype(np.float32)
    p = Params(Wi=u((C, D), D), bi=u((C,), D), Wf=u((C, C, D), C * D), bf=u((C,), C * D),
               W1=u((Q_DIM, D), D), b1=u((Q_DIM,), D))
    if nonlinear:
        p.W2, p.b2 = u((Q_DIM, Q_DIM), Q_DIM), u((Q_DIM,), Q_DIM)
    if passing_v:
        p.Wv, p.bv = u((D, D), D), u((D,), D)
    return p


def synthetic_bag(N: int, D: int, seed: int, kind: str = "uniform") -> np.ndarray:
    """Synthetic features: U[0,1) (post-ReLU/avg-pool embeddings are non-negative) or N(0,1)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((N, D), dtype=np.float32)
    return rng.standard_normal((N, D), dtype=np.float32)


# --------------------------------------------------------------------------- forward
def q_mlp(X: np.ndarray, p: Params) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """dsmil.py:31 (Linear-ReLU-Linear-Tanh) or :33 (Linear).  Returns (Q, H1)."""
    z1 = X @ p.W1.T + p.b1
    if not p.nonlinear:
        return z1, None
    h1 = np.maximum(z1, 0)
    return np.tanh(h1 @ p.W2.T + p.b2), h1


def critical_instances(c: np.ndarray) -> np.ndarray:
    """dsmil.py:52 keeps only row 0 of a descending sort == per-class arg-max.  Ties: the
    reference is implementation-defined; the product defines lowest index,
    and so does np.argmax.  NaN ranks first in torch.sort(descending); mirrored here."""
    c = np.asarray(c)
    key = np.where(np.isnan(c), np.inf, c)
    return np.argmax(key, axis=0).astype(np.int64)


@dataclass
class FwdOut:
    classes: np.ndarray      # [N, C]
    prediction_bag: np.ndarray  # [1, C]
    A: np.ndarray            # [N, C]
    B: np.ndarray            # [1, C, D]
    idx: np.ndarray          # [C] int64 critical instances
    Q: np.ndarray = field(repr=False, default=None)
    H1: Optional[np.ndarray] = field(repr=False, default=None)
    V: np.ndarray = field(repr=False, default=None)
    logits: np.ndarray = field(repr=False, default=None)  # [N, C] pre-softmax


def forward(X: np.ndarray, p: Params, dtype=np.float64, idx: Optional[np.ndarray] = None,
            v_mask: Optional[np.ndarray] = None) -> FwdOut:
    """MILNet.forward (dsmil.py:70-74) with FCLayer instance stream.

    ``idx`` lets a caller force the critical rows (used for tie cases).  ``v_mask`` is the
    (already scaled) dropout mask applied to X inside ``v`` when passing_v (dsmil.py:36).
    """
    X = np.asarray(X, dtype=dtype)
    p = p.astype(dtype)
    scale = dtype(SCALE_F32) if dtype is not np.float64 else np.float64(SCALE_F32)
    c = X @ p.Wi.T + p.bi                                   # dsmil.py:11
    if p.passing_v:                                          # dsmil.py:48
        Xv = X if v_mask is None else X * np.asarray(v_mask, dtype=dtype)
        V = np.maximum(Xv @ p.Wv.T + p.bv, 0)
    else:
        V = X
    Q, H1 = q_mlp(X, p)                                      # dsmil.py:49
    if idx is None:
        idx = critical_instances(c)                          # dsmil.py:52
    q_max, _ = q_mlp(X[idx], p)                              # dsmil.py:53-54
    L = (Q @ q_max.T) / scale                                # dsmil.py:55-56 (a division)
    e = np.exp(L - L.max(axis=0, keepdims=True))
    A = e / e.sum(axis=0, keepdims=True)                     # softmax over the instance axis
    Bm = A.T @ V                                             # dsmil.py:57
    C_, D_ = p.C, Bm.shape[1]
    pred = p.Wf.reshape(C_, C_ * D_) @ Bm.reshape(C_ * D_) + p.bf   # dsmil.py:44,59-61 Conv1d == GEMV
    return FwdOut(c, pred.reshape(1, C_), A, Bm.reshape(1, C_, D_), np.asarray(idx, dtype=np.int64),
                  Q=Q, H1=H1, V=V, logits=L)


# --------------------------------------------------------------------------- backward
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def caller_loss_grads(out: FwdOut, y: np.ndarray, pos_weight: Optional[float] = None):
    """0.5*BCEWithLogits(bag) + 0.5*BCEWithLogits(max instance) (train_tcga.py:67-71,
    train_mil.py:50-55).  Returns (loss, d_classes[N,C], d_pred[1,C])."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    C = y.size
    pw = 1.0 if pos_weight is None else float(pos_weight)
    pb = out.prediction_bag.reshape(-1).astype(np.float64)
    mx_idx = critical_instances(out.classes)
    pm = out.classes[mx_idx, np.arange(C)].astype(np.float64)

    def bce(z):
        # -[pw*y*log(sig z) + (1-y)*log(1-sig z)], mean over C
        ls = -np.logaddexp(0, -z)
        l1s = -np.logaddexp(0, z)
        return float(np.mean(-(pw * y * ls + (1 - y) * l1s)))

    def dbce(z):
        s = _sigmoid(z)
        return (-(pw * y * (1 - s)) + (1 - y) * s) / C

    loss = 0.5 * bce(pb) + 0.5 * bce(pm)
    d_pred = (0.5 * dbce(pb)).reshape(1, C)
    d_cls = np.zeros_like(out.classes, dtype=np.float64)
    d_cls[mx_idx, np.arange(C)] = 0.5 * dbce(pm)
    return loss, d_cls, d_pred


def backward(X: np.ndarray, p: Params, out: FwdOut, d_classes: Optional[np.ndarray],
             d_pred: Optional[np.ndarray], d_A: Optional[np.ndarray] = None,
             d_B: Optional[np.ndarray] = None, need_dX: bool = False,
             v_mask: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Manual reverse pass of ``forward`` in fp64 (checked against autograd through the
    reference in tests/test_oracle.py).  The arg-max indices are non-differentiable."""
    f = np.float64
    X = np.asarray(X, f)
    p = p.astype(f)
    C, D = p.C, p.D
    N = X.shape[0]
    A, Q, H1, V = (np.asarray(t, f) if t is not None else None for t in (out.A, out.Q, out.H1, out.V))
    Bm = np.asarray(out.B, f).reshape(C, -1)
    idx = out.idx
    g: Dict[str, np.ndarray] = {}
    dX = np.zeros_like(X) if need_dX else None

    dc = np.zeros((N, C)) if d_classes is None else np.asarray(d_classes, f)
    g["Wi"] = dc.T @ X
    g["bi"] = dc.sum(0)
    if need_dX:
        dX += dc @ p.Wi

    dp = np.zeros(C) if d_pred is None else np.asarray(d_pred, f).reshape(C)
    g["Wf"] = np.outer(dp, Bm.reshape(-1)).reshape(C, C, -1)
    g["bf"] = dp.copy()
    dB = (p.Wf.reshape(C, -1).T @ dp).reshape(C, -1)
    if d_B is not None:
        dB = dB + np.asarray(d_B, f).reshape(C, -1)

    dA = V @ dB.T                                           # [N, C]
    if d_A is not None:
        dA = dA + np.asarray(d_A, f)
    dV = A @ dB                                             # [N, Dv]
    dL = A * (dA - (A * dA).sum(0, keepdims=True)) / f(SCALE_F32)
    q_max = Q[idx]
    dQ = dL @ q_max                                         # [N,128]
    dqm = dL.T @ Q                                          # [C,128]
    np.add.at(dQ, idx, dqm)                                 # q_max = q(X[idx]) == Q[idx] rows

    if p.nonlinear:
        dz2 = dQ * (1 - Q * Q)
        g["W2"] = dz2.T @ H1
        g["b2"] = dz2.sum(0)
        dz1 = (dz2 @ p.W2) * (H1 > 0)
    else:
        dz1 = dQ
    g["W1"] = dz1.T @ X
    g["b1"] = dz1.sum(0)
    if need_dX:
        dX += dz1 @ p.W1

    if p.passing_v:
        Xv = X if v_mask is None else X * np.asarray(v_mask, f)
        dzv = dV * (V > 0)
        g["Wv"] = dzv.T @ Xv
        g["bv"] = dzv.sum(0)
        if need_dX:
            t = dzv @ p.Wv
            dX += t if v_mask is None else t * np.asarray(v_mask, f)
    elif need_dX:
        dX += dV
    if need_dX:
        g["X"] = dX
    return g


# --------------------------------------------------------------------------- sharded algebra
def shard_bounds(N: int, G: int) -> Sequence[Tuple[int, int]]:
    """Contiguous row blocks; first N % G ranks get one extra row."""
    base, rem = divmod(N, G)
    out, lo = [], 0
    for r in range(G):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def forward_sharded(X: np.ndarray, p: Params, G: int, dtype=np.float64) -> FwdOut:
    """Row-sharded forward: exchange 1 = (score, global idx, q row) candidates; exchange 2 =
    (m, s, partial B).  Must reproduce ``forward`` (tests assert it)."""
    X = np.asarray(X, dtype)
    p = p.astype(dtype)
    C = p.C
    scale = dtype(SCALE_F32)
    bounds = shard_bounds(X.shape[0], G)
    loc = []
    for lo, hi in bounds:
        Xr = X[lo:hi]
        if hi == lo:
            loc.append(None)
            continue
        c = Xr @ p.Wi.T + p.bi
        Q, _ = q_mlp(Xr, p)
        V = np.maximum(Xr @ p.Wv.T + p.bv, 0) if p.passing_v else Xr
        li = critical_instances(c)
        loc.append(dict(c=c, Q=Q, V=V, val=c[li, np.arange(C)], gidx=li + lo, qrow=Q[li, :]))
    # exchange 1: max score, lowest global index on ties
    q_max = np.zeros((C, Q_DIM), dtype)
    gidx = np.zeros(C, np.int64)
    for k in range(C):
        best = None
        for r, d in enumerate(loc):
            if d is None:
                continue
            v = d["val"][k]
            v = np.inf if np.isnan(v) else v
            cand = (-v, d["gidx"][k], r)
            if best is None or cand < best:
                best = cand
        gidx[k] = best[1]
        q_max[k] = loc[best[2]]["qrow"][k]
    # local attention partials
    for d in loc:
        if d is None:
            continue
        L = (d["Q"] @ q_max.T) / scale
        d["L"] = L
        d["m"] = L.max(0)
        e = np.exp(L - d["m"])
        d["s"] = e.sum(0)
        d["Bp"] = e.T @ d["V"]
    live = [d for d in loc if d is not None]
    M = np.max(np.stack([d["m"] for d in live]), axis=0)
    S = sum(d["s"] * np.exp(d["m"] - M) for d in live)
    Bm = sum(d["Bp"] * np.exp(d["m"] - M)[:, None] for d in live) / S[:, None]
    A = np.concatenate([np.exp(d["L"] - M) / S for d in live], 0)
    c_all = np.concatenate([d["c"] for d in live], 0)
    Dv = Bm.shape[1]
    pred = p.Wf.reshape(C, C * Dv) @ Bm.reshape(-1) + p.bf
    return FwdOut(c_all, pred.reshape(1, C), A, Bm.reshape(1, C, Dv), gidx)


def backward_sharded(X: np.ndarray, p: Params, out: FwdOut, d_classes: Optional[np.ndarray],
                     d_pred: Optional[np.ndarray], G: int) -> Dict[str, np.ndarray]:
    """Row-sharded reverse pass (SURVEY A.2 "Sharded", §8e): rank r holds rows [lo_r, hi_r) of X, classes, Q,
    H1, A and the replicated B / q_max / critical indices from the sharded forward.  Three reductions:
        reduce 1: t[k] = sum_n A[n,k] dA[n,k]                 (C scalars; the softmax-over-instances backward)
        reduce 2: dq_max[k,:] = sum_n dL[n,k] Q[n,:]          (C x 128; applied on the critical row's owner)
        reduce 3: parameter gradients                         (sum over ranks; Wf/bf are replicated, not summed)
    Must reproduce ``backward`` (identity v only; tests assert it)."""
    if p.passing_v:
        raise NotImplementedError("sharded backward: identity v only")
    f = np.float64
    X = np.asarray(X, f)
    p = p.astype(f)
    C = p.C
    A, Q, H1 = np.asarray(out.A, f), np.asarray(out.Q, f), (np.asarray(out.H1, f) if out.H1 is not None else None)
    Bm = np.asarray(out.B, f).reshape(C, -1)
    idx = out.idx
    q_max = Q[idx]                                            # replicated by the forward's exchange 1
    bounds = shard_bounds(X.shape[0], G)
    dp = np.zeros(C) if d_pred is None else np.asarray(d_pred, f).reshape(C)
    dc = np.zeros((X.shape[0], C)) if d_classes is None else np.asarray(d_classes, f)
    # replicated on every rank: bag classifier and dB
    g: Dict[str, np.ndarray] = {"Wf": np.outer(dp, Bm.reshape(-1)).reshape(C, C, -1), "bf": dp.copy()}
    dB = (p.Wf.reshape(C, -1).T @ dp).reshape(C, -1)
    # local part 1 + reduce 1
    dA = [X[lo:hi] @ dB.T for lo, hi in bounds]
    t = sum((A[lo:hi] * dA_r).sum(0) for (lo, hi), dA_r in zip(bounds, dA))
    # local part 2 + reduce 2
    dL = [A[lo:hi] * (dA_r - t) / f(SCALE_F32) for (lo, hi), dA_r in zip(bounds, dA)]
    dqm = sum(dL_r.T @ Q[lo:hi] for (lo, hi), dL_r in zip(bounds, dL))
    # local part 3 + reduce 3
    names = ["Wi", "bi", "W1", "b1"] + (["W2", "b2"] if p.nonlinear else [])
    for n in names:
        g[n] = np.zeros_like(getattr(p, n))
    for (lo, hi), dL_r in zip(bounds, dL):
        Xr, Qr = X[lo:hi], Q[lo:hi]
        dQ = dL_r @ q_max
        for k in range(C):
            if lo <= idx[k] < hi:                             # the owner of critical row k adds dq_max[k]
                dQ[idx[k] - lo] += dqm[k]
        if p.nonlinear:
            dz2 = dQ * (1 - Qr * Qr)
            g["W2"] += dz2.T @ H1[lo:hi]
            g["b2"] += dz2.sum(0)
            dz1 = (dz2 @ p.W2) * (H1[lo:hi] > 0)
        else:
            dz1 = dQ
        g["W1"] += dz1.T @ Xr
        g["b1"] += dz1.sum(0)
        g["Wi"] += dc[lo:hi].T @ Xr
        g["bi"] += dc[lo:hi].sum(0)
    return g


# --------------------------------------------------------------------------- split-precision emulation
def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (finite inputs)."""
    u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split_bf16(a: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    hi = bf16_round(a)
    lo = bf16_round(np.asarray(a, np.float32) - hi)
    return hi, lo


def matmul_3xbf16(X: np.ndarray, W: np.ndarray) -> np.ndarray:
    """Emulates the tensor-core path: X·Wᵀ ≈ Xh·Whᵀ + Xl·Whᵀ + Xh·Wlᵀ with exact products and
    (here fp64, on the device fp32) accumulation.  Used to predict the device path's error."""
    xh, xl = split_bf16(X)
    wh, wl = split_bf16(W)
    f = np.float64
    return (xh.astype(f) @ wh.astype(f).T + xl.astype(f) @ wh.astype(f).T + xh.astype(f) @ wl.astype(f).T)


# --------------------------------------------------------------------------- torch-CPU port (cpu_baseline kind="port")
class TorchPort:
    """The reference's op sequence (dsmil.py:10-12,46-62) on torch-CPU fp32 tensors, all host
    threads.  Functional (no nn.Module); timed by bench.py as the CPU baseline."""

    def __init__(self, p: Params, threads: Optional[int] = None):
        import os
        import torch
        self.torch = torch
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        self.Wi, self.bi, self.Wf, self.bf = t(p.Wi), t(p.bi), t(p.Wf), t(p.bf)
        self.W1, self.b1, self.W2, self.b2 = t(p.W1), t(p.b1), t(p.W2), t(p.b2)
        self.Wv, self.bv = t(p.Wv), t(p.bv)
        self.C, self.D = p.C, p.D

    def _q(self, x):
        F = self.torch.nn.functional
        z = F.linear(x, self.W1, self.b1)
        if self.W2 is None:
            return z
        return self.torch.tanh(F.linear(self.torch.relu(z), self.W2, self.b2))

    def forward(self, feats):
        torch = self.torch
        F = torch.nn.functional
        with torch.no_grad():
            scores = F.linear(feats, self.Wi, self.bi)
            V = feats if self.Wv is None else torch.relu(F.linear(feats, self.Wv, self.bv))
            Q = self._q(feats)
            order = torch.sort(scores, 0, descending=True).indices       # dsmil.py:52 (full sort, as the reference)
            crit = feats.index_select(0, order[0])
            qm = self._q(crit)
            att = torch.softmax((Q @ qm.t()) / torch.sqrt(torch.tensor(float(Q_DIM), dtype=torch.float32)), 0)
            bag = att.t() @ V
            pred = F.conv1d(bag.unsqueeze(0), self.Wf, self.bf).view(1, -1)
        return scores, pred, att, bag.unsqueeze(0)
