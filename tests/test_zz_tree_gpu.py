"""Tree mode on the device: `compute_tree_feats` with the product `embed_bag` staging loop vs the
reference's loop shape (compute_feats.py:93-118: batched low patches, one batch-of-1 forward per high patch,
numpy fusion), same ResNet-18 + InstanceNorm embedders.  Runs after the parity suites on purpose."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fusion", "cat"])
def test_tree_loop_matches_reference_style_loop(tmp_path, mode):
    import torchvision.models as models
    import torchvision.transforms.functional as VF
    from PIL import Image
    import dsmil as mil
    from dsmil_wsi_b200.embed import compute_tree_feats, list_tree_patches
    from dsmil_wsi_b200.formats import read_bag_csv
    rng = np.random.default_rng(5)
    bag = tmp_path / "WSI" / "ds" / "pyramid" / "c0" / "slideT"
    bag.mkdir(parents=True)
    for i, nh in enumerate([3, 0, 5, 1]):
        Image.fromarray(rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)).save(bag / f"{i}_0.jpeg", quality=70)
        if nh:
            (bag / f"{i}_0").mkdir()
        for j in range(nh):
            Image.fromarray(rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)).save(bag / f"{i}_0" / f"{j}_1.jpeg",
                                                                                       quality=70)
    torch.backends.cudnn.allow_tf32 = False

    def embedder(seed):
        torch.manual_seed(seed)
        resnet = models.resnet18(weights=None, norm_layer=torch.nn.InstanceNorm2d)
        resnet.fc = torch.nn.Identity()
        return mil.IClassifier(resnet, 512, output_class=2).cuda().eval()
    low_net, high_net = embedder(1), embedder(2)
    got = {}
    args = argparse.Namespace(batch_size=4, num_workers=2, tree_fusion=mode)
    compute_tree_feats(args, [str(bag)], low_net, high_net, save_path=str(tmp_path / "out"),
                       sink=lambda d, f: got.update(feats=f.clone()))
    low, high = list_tree_patches(str(bag))
    with torch.no_grad():
        lf = torch.cat([low_net(torch.stack([VF.to_tensor(Image.open(p)) for p in low[i:i + 4]]).cuda())[0]
                        for i in range(0, len(low), 4)]).cpu().numpy()
        rows = []
        for idx, hl in enumerate(high):
            for hp in hl:
                f = high_net(VF.to_tensor(Image.open(hp)).float().cuda()[None, :])[0].cpu().numpy()
                rows.extend(f + 0.25 * lf[idx] if mode == "fusion" else np.concatenate((f, lf[idx][None, :]), axis=-1))
    ref = np.stack(rows)
    assert got["feats"].is_cuda and got["feats"].shape == ref.shape == (9, 512 if mode == "fusion" else 1024)
    # batch-of-1 vs batched convolutions may pick different cuDNN algorithms; a wrong parent or row order is O(1)
    assert np.abs(got["feats"].cpu().numpy() - ref).max() < 2e-3 * np.abs(ref).max()
    csv = read_bag_csv(os.path.join(str(tmp_path / "out"), "c0", "slideT.csv"))
    assert csv.shape == ref.shape and np.allclose(csv, got["feats"].cpu().numpy(), atol=6e-5)
