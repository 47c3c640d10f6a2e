"""Runs an UNMODIFIED reference caller script (oracle/_ref/train_tcga.py, train_mil.py) with either this repo's
`dsmil` shim or the reference's own `dsmil.py` resolving `import dsmil as mil` (train_tcga.py:224, train_mil.py:122).
Test infrastructure (tests/test_zz_acceptance_gpu.py); seeds every RNG the scripts draw from so that the two runs
see the same shuffles, initialisations and patch drop-outs.

    python tests/run_reference_caller.py {ours|ref} <script.py> [script args...]
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    which, script = sys.argv[1], sys.argv[2]
    sys.argv = [script] + sys.argv[3:]
    # runpy.run_path on a FILE does not put the script's directory on sys.path, so which `dsmil` wins is decided
    # here: the repo root (our shim) or oracle/_ref (the reference's module).
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, REF, os.path.join(ROOT, "tests"))]
    sys.path.insert(0, ROOT if which == "ours" else REF)
    import random
    import numpy as np
    import torch
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    torch.cuda.manual_seed_all(0)
    import dsmil
    print("DSMIL_MODULE=" + os.path.abspath(dsmil.__file__), flush=True)
    if which == "ours":
        from dsmil_wsi_b200 import _lib
        n0 = _lib.launch_count()
    runpy.run_path(os.path.join(REF, script), run_name="__main__")
    if which == "ours":
        print("\nDSMIL_LAUNCHES=%d" % (_lib.launch_count() - n0), flush=True)


if __name__ == "__main__":
    main()
