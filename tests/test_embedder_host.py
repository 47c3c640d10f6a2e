"""Host-side wiring of the embedder fusion (no GPU): fuse_instance_norm re-wires exactly the plain InstanceNorm2d
layers of a torchvision ResNet and leaves parameters / state_dict keys untouched."""
import pytest
import torch


def _resnet(name, norm):
    import torchvision.models as models
    m = getattr(models, name)(weights=None, norm_layer=norm)
    m.fc = torch.nn.Identity()
    return m


@pytest.mark.parametrize("name,expected", [("resnet18", 20), ("resnet34", 36), ("resnet50", 53)])
def test_fuse_counts_and_state_dict(name, expected):
    from dsmil_wsi_b200.embedder import fuse_instance_norm
    m = _resnet(name, torch.nn.InstanceNorm2d)              # compute_feats.py:146-170 (norm_layer='instance')
    keys = list(m.state_dict().keys())
    assert fuse_instance_norm(m) == expected
    assert fuse_instance_norm(m) == 0
    assert list(m.state_dict().keys()) == keys


def test_batchnorm_backbones_are_left_alone():
    from dsmil_wsi_b200.embedder import fuse_instance_norm
    m = _resnet("resnet18", torch.nn.BatchNorm2d)           # compute_feats.py norm_layer='batch'
    assert fuse_instance_norm(m) == 0
    y = m.eval()(torch.zeros(1, 3, 64, 64))                  # untouched forward still runs on CPU
    assert y.shape == (1, 512)


def test_fused_forward_refuses_cpu_tensors():
    from dsmil_wsi_b200.embedder import fuse_instance_norm
    m = _resnet("resnet18", torch.nn.InstanceNorm2d)
    for p in m.parameters():
        p.requires_grad = False
    fuse_instance_norm(m)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 3, 64, 64))                  # no CPU path
    for p in m.parameters():                                 # under autograd the original forward runs (CPU is fine there)
        p.requires_grad = True
    assert m(torch.zeros(1, 3, 64, 64)).shape == (1, 512)
