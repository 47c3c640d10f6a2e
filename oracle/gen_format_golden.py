"""Generate tests/golden/formats/* from the UNMODIFIED reference readers (test infrastructure).

Run in the authoring container only (needs /root/reference, read-only):

    python oracle/gen_format_golden.py

Writes small synthetic input files in the reference's on-disk formats and records what
the reference's own functions make of them:

    train_tcga.get_bag_feats      (train_tcga.py:19-34)   bag CSV -> features, label vector
    train_tcga.generate_pt_files  (train_tcga.py:36-51)   index CSV -> temp_train/*.pt
    train_mil.get_data / get_bag  (train_mil.py:17-40)    svm text -> instances, bags

The only intervention: `sklearn.utils.shuffle` inside those modules is replaced by the identity while the
fixtures are recorded, so rows come back in file order (the reference shuffles them with the unseeded global
numpy RNG; order carries no information for a bag).  The reference's code is otherwise imported as-is.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DSMIL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "formats")


def write_inputs():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    # bag feature CSVs exactly as compute_feats.py:80-82 writes them
    shapes = {"bag_a": (7, 5), "bag_b": (1, 5), "bag_c": (23, 5)}
    for name, (n, d) in shapes.items():
        feats = np.abs(rng.standard_normal((n, d))).astype(np.float32) * np.float32(1.7)
        feats[0, 0] = 0.0
        if n > 2:
            feats[2, 1] = 12345.678       # a value wider than the usual range
            feats[1, 2] = 0.00004         # rounds to 0.0000
        pd.DataFrame(feats).to_csv(os.path.join(OUT, name + ".csv"), index=False, float_format="%.4f")
    # dataset index as compute_feats.py:249-260 writes it (paths relative to the fixture directory)
    with open(os.path.join(OUT, "index.csv"), "w") as f:
        f.write("0,label\nbag_a.csv,0\nbag_b.csv,1\nbag_c.csv,2\n")
    # classic-MIL svm text with the quirks train_mil.get_data has to survive
    lines = ["0:0:-1 0:9.0 1:9.0 2:9.0 3:9.0"]                     # swallowed as the CSV header
    inst = 1
    bags = [(-1, 3), (1, 2), (1, 1), (-1, 4)]
    for b, (lab, n) in enumerate(bags):
        for _ in range(n):
            v = rng.standard_normal(4)
            toks = [f"{j}:{v[j]:.6f}" for j in range(4)]
            if inst == 2:
                toks[1] = "7:0.5"              # feature index is ignored: lands in slot 1
            if inst == 3:
                toks[2] = "2:1e-3"             # exponent notation
            line = f"{inst}:{b}:{lab} " + " ".join(toks)
            if inst == 4:
                line += " "                    # trailing space -> an extra zero slot
            lines.append(line)
            inst += 1
    with open(os.path.join(OUT, "toy.svm"), "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    write_inputs()
    sys.path.insert(0, REF)
    import train_mil
    import train_tcga
    train_tcga.shuffle = lambda x: x            # see the module docstring
    gold = {}
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        index = pd.read_csv("index.csv")
        for C in (1, 2, 3, 4):
            args = argparse.Namespace(dataset="fixture", num_classes=C)
            for i in range(len(index)):
                label, feats, path = train_tcga.get_bag_feats(index.iloc[i], args)
                stem = os.path.splitext(path)[0]
                gold[f"label_C{C}_{stem}"] = np.asarray(label, dtype=np.float64)
                if C == 1:
                    gold[f"feats_{stem}"] = torch.tensor(np.array(feats), dtype=torch.float32).numpy()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            for n in ("index.csv", "bag_a.csv", "bag_b.csv", "bag_c.csv"):
                os.symlink(os.path.join(OUT, n), n)
            args = argparse.Namespace(dataset="fixture", num_classes=3)
            train_tcga.generate_pt_files(args, pd.read_csv("index.csv"))
            for n in ("bag_a", "bag_b", "bag_c"):
                gold[f"pt_C3_{n}"] = torch.load(os.path.join("temp_train", n + ".pt")).numpy()
            os.chdir(OUT)
        data = train_mil.get_data("toy.svm")
        gold["svm_ids"] = np.array([[d[0], d[1], d[2]] for d in data], dtype=np.int64)
        gold["svm_len"] = np.array([len(d[3]) for d in data], dtype=np.int64)
        gold["svm_vals"] = np.concatenate([d[3] for d in data]).astype(np.float64)
        num_bag = data[-1][1] + 1
        gold["svm_num_bag"] = np.int64(num_bag)
        for b in range(num_bag):
            bag = train_mil.get_bag(data, b)
            gold[f"svm_bag{b}_label"] = np.int64(bag[0, 2])
            gold[f"svm_bag{b}_n"] = np.int64(bag.shape[0])
        # fold bookkeeping of train_mil.py:99-110
        for n_items, fold in ((23, 5), (92, 10), (10, 10)):
            for index in range(fold):
                tr, te = train_mil.cross_validation_set(list(range(n_items)), fold, index)
                gold[f"cv_{n_items}_{fold}_{index}_train"] = np.array(tr, dtype=np.int64)
                gold[f"cv_{n_items}_{fold}_{index}_test"] = np.array(te, dtype=np.int64)
        labs = [-1, 1, 1, -1, -1, 0, 1, -1]
        gold["pos_weight_labels"] = np.array(labs, dtype=np.int64)
        gold["pos_weight"] = np.float64(train_mil.compute_pos_weight([[l, None] for l in labs]))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **gold)
    print(f"wrote {len(gold)} arrays to {OUT}/expected.npz")


if __name__ == "__main__":
    main()
