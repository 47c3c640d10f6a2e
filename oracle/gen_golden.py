"""Generate tests/golden/*.npz from the UNMODIFIED reference (test infrastructure).

Run in the authoring container only (needs /root/reference, read-only):

    python oracle/gen_golden.py

Imports /root/reference/dsmil.py as-is, runs MILNet(FCLayer, BClassifier) forward and the
callers' loss backward (train_tcga.py:67-72) on CPU fp32, and stores outputs + gradients.
Inputs are NOT stored: they are regenerated from seeds by oracle.dsmil_oracle.synthetic_bag /
random_params (a CRC of each regenerated array is stored to detect RNG drift).  The shipped
example weights (the reference's only known-answer artefacts) are stored as data.
"""
import importlib.util
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import dsmil_oracle as orc  # noqa: E402

REF = os.environ.get("DSMIL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("_ref_dsmil", os.path.join(REF, "dsmil.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def build_ref_model(ref, p: orc.Params, dropout_v=0.0):
    net = ref.MILNet(ref.FCLayer(p.D, p.C),
                     ref.BClassifier(p.D, p.C, dropout_v=dropout_v, nonlinear=p.nonlinear, passing_v=p.passing_v))
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float32))
    sd = {"i_classifier.fc.0.weight": t(p.Wi), "i_classifier.fc.0.bias": t(p.bi),
          "b_classifier.fcc.weight": t(p.Wf), "b_classifier.fcc.bias": t(p.bf)}
    if p.nonlinear:
        sd.update({"b_classifier.q.0.weight": t(p.W1), "b_classifier.q.0.bias": t(p.b1),
                   "b_classifier.q.2.weight": t(p.W2), "b_classifier.q.2.bias": t(p.b2)})
    else:
        sd.update({"b_classifier.q.weight": t(p.W1), "b_classifier.q.bias": t(p.b1)})
    if p.passing_v:
        sd.update({"b_classifier.v.1.weight": t(p.Wv), "b_classifier.v.1.bias": t(p.bv)})
    net.load_state_dict(sd, strict=True)
    return net.eval()  # eval: dropout in v is the only mode-dependent op (dsmil.py:36)


GRAD_KEYS = {
    "Wi": "i_classifier.fc.0.weight", "bi": "i_classifier.fc.0.bias",
    "Wf": "b_classifier.fcc.weight", "bf": "b_classifier.fcc.bias",
}


def run_case(ref, name, p: orc.Params, X: np.ndarray, y: np.ndarray, extra=None, with_dx=False):
    net = build_ref_model(ref, p)
    xt = torch.from_numpy(X).requires_grad_(with_dx)
    classes, pred, A, B = net(xt)
    # callers' loss, train_tcga.py:67-71
    crit = torch.nn.BCEWithLogitsLoss()
    yt = torch.from_numpy(y.astype(np.float32)).view(1, -1)
    mx, mxi = torch.max(classes, 0)
    loss = 0.5 * crit(pred.view(1, -1), yt) + 0.5 * crit(mx.view(1, -1), yt)
    loss.backward()
    out = dict(classes=classes.detach().numpy(), pred=pred.detach().numpy(), A=A.detach().numpy(),
               B=B.detach().numpy(), loss=np.float32(loss.item()), y=y.astype(np.float32),
               x_crc=crc(X), N=np.int64(X.shape[0]), D=np.int64(p.D), C=np.int64(p.C),
               nonlinear=np.int64(p.nonlinear), passing_v=np.int64(p.passing_v))
    # index actually used by the reference (row 0 of the descending sort, dsmil.py:52)
    with torch.no_grad():
        _, mi = torch.sort(classes, 0, descending=True)
    out["idx"] = mi[0].numpy().astype(np.int64)
    sd_names = dict(GRAD_KEYS)
    if p.nonlinear:
        sd_names.update(W1="b_classifier.q.0.weight", b1="b_classifier.q.0.bias",
                        W2="b_classifier.q.2.weight", b2="b_classifier.q.2.bias")
    else:
        sd_names.update(W1="b_classifier.q.weight", b1="b_classifier.q.bias")
    if p.passing_v:
        sd_names.update(Wv="b_classifier.v.1.weight", bv="b_classifier.v.1.bias")
    named = dict(net.named_parameters())
    for short, full in sd_names.items():
        out["g_" + short] = named[full].grad.numpy().astype(np.float32)
    if with_dx:
        out["g_X"] = xt.grad.numpy().astype(np.float32)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: N={X.shape[0]} D={p.D} C={p.C} loss={loss.item():.6f} idx={out['idx']}")


def tie_free(X, p, gap=1e-4):
    c = X.astype(np.float64) @ p.Wi.astype(np.float64).T + p.bi
    if c.shape[0] < 2:
        return True
    s = np.sort(c, axis=0)
    return bool(np.all(s[-1] - s[-2] > gap))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for the stored fp32 outputs

    # (1),(2): shipped checkpoints = the reference's known-answer artefacts
    for tag, N, seed, y in (("tcga", 1500, 0, [1, 0]), ("c16", 1201, 1, [1])):
        sd = torch.load(os.path.join(REF, "example_aggregator_weights", f"{tag}_aggregator.pth"), map_location="cpu")
        p = orc.params_from_state_dict(sd)
        X = orc.synthetic_bag(N, p.D, seed, "uniform")
        assert tie_free(X, p)
        w = {"w_" + k: getattr(p, k) for k in ("Wi", "bi", "Wf", "bf", "W1", "b1", "W2", "b2")}
        run_case(ref, f"shipped_{tag}", p, X, np.array(y), extra=w)

    # (3..): seeded random weights (regenerated by oracle.random_params in the tests)
    cases = [
        # name,            D,   C, N,    xseed, wseed, kind,     nonlinear, passing_v, y,        wscale, dx
        ("rand_d512_c2",   512, 2, 777,  10,    110,   "normal", True,  False, [0, 1],     1.0, False),
        ("rand_d512_c1",   512, 1, 2049, 11,    111,   "uniform", True, False, [0],        3.0, False),
        ("lin_d512_c3",    512, 3, 300,  12,    112,   "normal", False, False, [1, 0, 0],  1.0, False),
        ("musk_d166_n7",   166, 1, 7,    13,    113,   "normal", True,  False, [1],        1.0, True),
        ("musk_d166_n1",   166, 1, 1,    14,    114,   "normal", True,  False, [0],        1.0, True),
        ("musk_d166_n2",   166, 1, 2,    15,    115,   "normal", True,  False, [1],        1.0, False),
        ("eleph_d230_n33", 230, 1, 33,   16,    116,   "normal", True,  False, [1],        1.0, False),
        ("pv_d96_c2",      96,  2, 150,  17,    117,   "normal", True,  True,  [1, 1],     1.0, True),
        ("pvlin_d64_c5",   64,  5, 90,   18,    118,   "normal", False, True,  [0, 1, 0, 0, 1], 1.0, False),
        ("tree_d1024_c2",  1024, 2, 260, 19,    119,   "uniform", True, False, [1, 0],     2.0, False),
    ]
    for (name, D, C, N, xs, ws, kind, nl, pv, y, wscale, dx) in cases:
        p = orc.random_params(D, C, ws, nonlinear=nl, passing_v=pv, scale=wscale)
        X = orc.synthetic_bag(N, D, xs, kind)
        assert tie_free(X, p), name
        meta = dict(xseed=np.int64(xs), wseed=np.int64(ws), kind=np.array(kind), wscale=np.float64(wscale),
                    w_crc=crc(p.W1))
        run_case(ref, name, p, X, np.array(y), extra=meta, with_dx=dx)


if __name__ == "__main__":
    main()
