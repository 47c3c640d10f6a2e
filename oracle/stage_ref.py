"""Stages the UNMODIFIED reference Python sources of the hot path into ``oracle/_ref/`` (git-ignored,
NOT gpurun-ignored) so that they travel to the GPU box, where ``/root/reference`` does not exist.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing under ``dsmil_wsi_b200/`` reads ``oracle/_ref``.  Users:
  * ``bench.py --impl reference`` and the ``torch_eager_gpu`` leg: the reference's own ``MILNet`` timed on the
    host cores / through PyTorch eager on the GPU (``cpu_baseline.kind == "reference"``);
  * ``bench.py``'s ``embed_from_files`` leg: the reference's own ``compute_feats.compute_feats`` loop (DataLoader
    workers + PIL + ``.float().cuda()`` + pandas CSV) run on a folder of patch files;
  * ``tests/test_zz_acceptance_gpu.py``: the unmodified ``train_tcga.py`` / ``train_mil.py`` run with THIS repo's
    ``dsmil.py`` shim ahead of them on ``sys.path`` (the callers are the acceptance harness).

The files are byte-for-byte copies made at build time by ``__graft_entry__.build()`` when the reference checkout
is present; they are never committed (``.gitignore`` lists ``oracle/_ref/``) and never edited.  A manifest with
the SHA-256 of every staged file is written next to them so that a test can check they are unmodified.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DSMIL_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("dsmil.py", "train_tcga.py", "train_mil.py", "compute_feats.py")


def staged(name: str = "dsmil.py"):
    """Path of a staged reference file, or None when the reference was not staged (fresh clone without
    /root/reference): callers fall back to the oracle port and say so."""
    p = os.path.join(REF_DST, name)
    return p if os.path.exists(p) else None


def stage(force: bool = False):
    """Copies the reference files (only when /root/reference exists).  Returns the list of staged paths."""
    if not os.path.isdir(REF_SRC):
        return [p for p in (staged(f) for f in FILES) if p]
    os.makedirs(REF_DST, exist_ok=True)
    manifest = {}
    out = []
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        if not os.path.exists(src):
            continue
        data = open(src, "rb").read()
        manifest[f] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
        out.append(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "sha256": manifest}, fh, indent=1)
    return out


def load_reference_dsmil():
    """Imports the staged, unmodified reference ``dsmil.py`` under the module name ``_ref_dsmil`` (the top-level
    name ``dsmil`` is this repo's shim).  Returns the module or None."""
    import importlib.util
    path = staged("dsmil.py")
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("_ref_dsmil", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod




def load_reference_compute_feats():
    """Imports the staged, unmodified reference ``compute_feats.py`` as ``_ref_compute_feats`` with ITS ``dsmil``
    (the staged reference module) bound to the name it imports.  Returns the module or None."""
    import importlib.util
    import sys
    path = staged("compute_feats.py")
    ref_dsmil = load_reference_dsmil()
    if path is None or ref_dsmil is None:
        return None
    if "_ref_compute_feats" in sys.modules:
        return sys.modules["_ref_compute_feats"]
    saved = sys.modules.get("dsmil")
    sys.modules["dsmil"] = ref_dsmil
    try:
        spec = importlib.util.spec_from_file_location("_ref_compute_feats", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_ref_compute_feats"] = mod
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["dsmil"] = saved
        else:
            sys.modules.pop("dsmil", None)
    return mod


if __name__ == "__main__":
    print("\n".join(stage()))
