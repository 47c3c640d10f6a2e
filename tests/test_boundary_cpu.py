"""Host-side contract of the drop-in boundary -- runs without a GPU.
Checks module identity / constructor signatures / state_dict layout against the reference's
shipped checkpoints (stored in tests/golden), and that the product refuses to run without CUDA
instead of falling back."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import ROOT, load_golden


def test_import_as_reference_module_name():
    import dsmil as mil
    for n in ("FCLayer", "IClassifier", "BClassifier", "MILNet"):
        assert hasattr(mil, n)


def _sd_from_golden(name):
    g, p, _ = load_golden(name)
    t = lambda a: torch.from_numpy(np.array(a))
    return {"i_classifier.fc.0.weight": t(p.Wi), "i_classifier.fc.0.bias": t(p.bi),
            "b_classifier.q.0.weight": t(p.W1), "b_classifier.q.0.bias": t(p.b1),
            "b_classifier.q.2.weight": t(p.W2), "b_classifier.q.2.bias": t(p.b2),
            "b_classifier.fcc.weight": t(p.Wf), "b_classifier.fcc.bias": t(p.bf)}


@pytest.mark.parametrize("name,C", [("shipped_tcga", 2), ("shipped_c16", 1)])
def test_shipped_checkpoints_load_strict(name, C):
    import dsmil as mil
    net = mil.MILNet(mil.FCLayer(in_size=512, out_size=C), mil.BClassifier(input_size=512, output_class=C))
    missing = net.load_state_dict(_sd_from_golden(name), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(net.state_dict().keys()) == list(_sd_from_golden(name).keys())  # same order as the reference


def test_variant_state_dict_keys():
    import dsmil as mil
    b = mil.BClassifier(64, 3, dropout_v=0.2, nonlinear=False, passing_v=True)
    assert sorted(b.state_dict().keys()) == ["fcc.bias", "fcc.weight", "q.bias", "q.weight", "v.1.bias", "v.1.weight"]
    assert b.fcc.weight.shape == (3, 3, 64) and isinstance(b.fcc, nn.Conv1d)
    ic = mil.IClassifier(nn.Flatten(), 32, 2)
    assert sorted(ic.state_dict().keys()) == ["fc.bias", "fc.weight"]
    # `--non_linearity` arrives as a float (train_tcga.py:213,237): truthy -> nonlinear
    assert isinstance(mil.BClassifier(16, 1, nonlinear=1.0).q, nn.Sequential)
    assert isinstance(mil.BClassifier(16, 1, nonlinear=0.0).q, nn.Linear)


def test_caller_idioms_work_on_the_module_tree():
    """train_tcga.py:229-243 (apply+isinstance init, Adam over parameters), :389-390 (deepcopy of .cpu()),
    testing_tcga.py:141-144 (re-keying FCLayer.fc.0 -> IClassifier.fc and sub-module reassignment)."""
    import dsmil as mil
    net = mil.MILNet(mil.FCLayer(32, 2), mil.BClassifier(32, 2))
    seen = []

    def init(m):
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv1d)):
            nn.init.orthogonal_(m.weight)
            nn.init.constant_(m.bias, 0)
            seen.append(type(m).__name__)
    net.apply(init)
    assert seen.count("Linear") == 3 and seen.count("Conv1d") == 1
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.5, 0.9), weight_decay=1e-3)
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == 2 * 32 + 2 + 128 * 32 + 128 + 128 * 128 + 128 + 2 * 2 * 32 + 2
    clone = copy.deepcopy(net.cpu())
    assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), net.state_dict().values()))
    sd = net.state_dict()
    ic = mil.IClassifier(nn.Identity(), 32, 2)
    ic.load_state_dict({"fc.weight": sd["i_classifier.fc.0.weight"], "fc.bias": sd["i_classifier.fc.0.bias"]})
    net.i_classifier = ic
    assert isinstance(net.i_classifier, mil.IClassifier)
    net.train(); assert net.training
    net.eval(); assert not net.b_classifier.training


def test_no_cpu_fallback():
    import dsmil as mil
    net = mil.MILNet(mil.FCLayer(16, 1), mil.BClassifier(16, 1))
    with pytest.raises(RuntimeError, match="CUDA only"):
        net(torch.randn(5, 16))
    with pytest.raises(RuntimeError, match="CUDA only"):
        net.i_classifier(torch.randn(5, 16))
    with pytest.raises(RuntimeError, match="CUDA only"):
        net.b_classifier(torch.randn(5, 16), torch.randn(5, 1))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dsmil_wsi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "dsmil_oracle" not in src, f


def test_library_exports_every_declared_symbol():
    from dsmil_wsi_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "dsmil_b200.h")).read()
    declared = set(re.findall(r"\b(dsmil_[a-z0-9_]+)\s*\(", header))
    declared -= {"dsmil_status", "dsmil_params", "dsmil_grads"}
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dsmil_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.dsmil_abi_version() == 1
    # pure-host helpers are callable without a GPU
    assert lib.dsmil_cand_floats(2) == 264 and lib.dsmil_cand_floats(1) == 132   # 131*C padded to 4 floats
    assert lib.dsmil_rec_floats(2, 512) == 2 * (2 + 512) and lib.dsmil_rec_floats(1, 166) == 168
    P = _lib.DsmilParams(512, 2, 1, 0)
    assert lib.dsmil_forward_workspace_bytes(ctypes.byref(P), 10000) > 2 * 10000 * 128 * 4
    assert lib.dsmil_backward_workspace_bytes(ctypes.byref(P), 10000, 0) > 0
    bad = _lib.DsmilParams(512, 99, 1, 0)
    assert lib.dsmil_forward_workspace_bytes(ctypes.byref(bad), 10) == 0


def test_host_library_exports_every_declared_symbol():
    """include/dsmil_host.h <-> libdsmil_host.so <-> _hostlib.SIGNATURES."""
    import re
    from dsmil_wsi_b200 import _hostlib
    hdr = open(os.path.join(ROOT, "include", "dsmil_host.h")).read()
    declared = set(re.findall(r"\b(dsmil_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_hostlib.SIGNATURES), declared ^ set(_hostlib.SIGNATURES)
    lib = _hostlib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dsmil_host_abi_version() == 1 and "#define DSMIL_HOST_ABI_VERSION 1" in hdr
