"""Generate tests/golden/tree/* from the UNMODIFIED reference tree-mode loop (test infrastructure).

Run in the authoring container only (needs /root/reference, read-only):

    python oracle/gen_tree_golden.py

Builds a tiny two-magnification bag of synthetic JPEG patches in the reference's folder layout
(`<bag>/<x>.jpeg` low patches, `<bag>/<x>/<y>.jpeg` high patches, compute_feats.py:91,101-103), runs
`compute_feats.compute_tree_feats` (compute_feats.py:84-126) on CPU for both `--tree_fusion` modes with two
small seeded `dsmil.IClassifier` embedders, and records the CSVs it wrote, the traversal it used, and the
embedder weights.  The only intervention: `torch.Tensor.cuda` is the identity while the reference runs (its
loop hard-codes `.cuda()`, there is no GPU here).
"""
import argparse
import glob
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch
import torch.nn as nn
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DSMIL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "tree")
BAG = os.path.join("WSI", "ds", "pyramid", "c0", "slideT")
FEATS = 8


def backbone(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, FEATS, 3, stride=2), nn.InstanceNorm2d(FEATS), nn.ReLU(),
                         nn.AdaptiveAvgPool2d(1), nn.Flatten())


def write_patches():
    rng = np.random.default_rng(11)
    bag = os.path.join(OUT, BAG)
    layout = {"0_0.jpeg": ["0_0.jpeg", "0_1.jpeg"], "0_1.jpeg": [], "1_0.jpg": ["2_0.jpeg", "2_1.jpg", "3_0.jpeg"]}
    for low, highs in layout.items():
        os.makedirs(bag, exist_ok=True)
        Image.fromarray(rng.integers(0, 256, size=(16, 16, 3), dtype=np.uint8)).save(os.path.join(bag, low), quality=80)
        folder = os.path.join(bag, os.path.splitext(low)[0])
        for h in highs:
            os.makedirs(folder, exist_ok=True)
            Image.fromarray(rng.integers(0, 256, size=(16, 16, 3), dtype=np.uint8)).save(os.path.join(folder, h), quality=80)


def main():
    write_patches()
    sys.path.insert(0, REF)
    import compute_feats as ref
    import dsmil as ref_mil
    torch.Tensor.cuda = lambda self, *a, **k: self
    low = ref_mil.IClassifier(backbone(1), FEATS, output_class=2)
    high = ref_mil.IClassifier(backbone(2), FEATS, output_class=2)
    gold = {f"low.{k}": v.numpy() for k, v in low.state_dict().items()}
    gold.update({f"high.{k}": v.numpy() for k, v in high.state_dict().items()})
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        # the traversal the reference is about to use (same globs, same process => same order)
        lows = glob.glob(os.path.join(BAG, "*.jpg")) + glob.glob(os.path.join(BAG, "*.jpeg"))
        rows = []
        for lp in lows:
            folder = os.path.dirname(lp) + os.sep + os.path.splitext(os.path.basename(lp))[0]
            for hp in glob.glob(folder + os.sep + "*.jpg") + glob.glob(folder + os.sep + "*.jpeg"):
                rows.append((os.path.basename(lp), os.path.basename(hp)))
        gold["row_low"] = np.array([r[0] for r in rows])
        gold["row_high"] = np.array([r[1] for r in rows])
        for mode in ("fusion", "cat"):
            with tempfile.TemporaryDirectory() as tmp:
                args = argparse.Namespace(batch_size=2, num_workers=0, tree_fusion=mode)
                ref.compute_tree_feats(args, [BAG], low, high, save_path=tmp)
                csv = os.path.join(tmp, "c0", "slideT.csv")
                gold[f"csv_{mode}"] = np.array(open(csv).read())
                gold[f"feats_{mode}"] = pd.read_csv(csv).to_numpy()
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **gold)
    print("rows:", rows)
    print({k: v.shape for k, v in gold.items() if k.startswith("feats")})


if __name__ == "__main__":
    main()
