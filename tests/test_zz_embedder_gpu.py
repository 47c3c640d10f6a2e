"""Embedder kernel: fused InstanceNorm (+ residual) (+ ReLU) against torch's own ops, and the re-wired
torchvision ResNet-18 of compute_feats.py:146-170 against the untouched module."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, res, relu, eps=1e-5):
    y = torch.nn.functional.instance_norm(x, eps=eps)        # affine=False, no running stats: nn.InstanceNorm2d defaults
    if res is not None:
        y = y + res
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("shape", [(4, 64, 112, 112), (3, 64, 56, 56), (2, 128, 28, 28), (2, 256, 14, 14), (2, 512, 7, 7),
                                   (2, 3, 5, 7), (1, 3, 37, 31), (1, 2, 128, 128), (2, 5, 1, 2)])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (False, False)])
def test_instnorm_act_matches_torch(shape, with_res, relu):
    from dsmil_wsi_b200.embedder import instnorm_act
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g, device="cuda") * 3.0 + 1.5
    res = torch.randn(*shape, generator=g, device="cuda") if with_res else None
    want = _ref(x, res, relu)
    got = instnorm_act(x, res, relu)
    scale = float(want.abs().max().clamp_min(1.0))
    assert float((got - want).abs().max()) <= 2e-5 * scale, shape
    # in place
    x2 = x.clone()
    out = instnorm_act(x2, res, relu, out=x2)
    assert out.data_ptr() == x2.data_ptr() and torch.equal(out, got)


def test_instnorm_act_rejects_cpu_and_bad_shapes():
    from dsmil_wsi_b200.embedder import instnorm_act
    with pytest.raises(RuntimeError):
        instnorm_act(torch.zeros(1, 1, 4, 4))
    with pytest.raises(TypeError):
        instnorm_act(torch.zeros(4, 4, device="cuda"))
    with pytest.raises(RuntimeError):                        # one plane larger than the staged maximum (16 384 elements)
        instnorm_act(torch.zeros(1, 1, 129, 128, device="cuda"))


def test_fused_resnet18_matches_untouched_module():
    import copy
    import torchvision.models as models
    from dsmil_wsi_b200.embedder import fuse_instance_norm
    import dsmil as mil
    torch.manual_seed(0)
    resnet = models.resnet18(weights=None, norm_layer=torch.nn.InstanceNorm2d)    # compute_feats.py:154
    resnet.fc = torch.nn.Identity()
    for p in resnet.parameters():
        p.requires_grad = False
    ref = copy.deepcopy(resnet).cuda().eval()
    fused = resnet.cuda().eval()
    keys_before = list(fused.state_dict().keys())
    n = fuse_instance_norm(fused)
    assert n == 20                                            # stem 1 + 8 blocks x 2 + 3 downsample norms
    assert fuse_instance_norm(fused) == 0                     # idempotent
    assert list(fused.state_dict().keys()) == keys_before     # checkpoints of the reference still load
    x = torch.rand(6, 3, 224, 224, device="cuda")
    # full-fp32 convolutions for the comparison: with TF32 (torch's default for cuDNN) a 1e-6 difference after one norm
    # flips 10-bit roundings in the next convolution and the two arms drift apart by ~1e-3 -- of either arm's own noise
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            a = ref(x)
            b = fused(x)
            ic = mil.IClassifier(fused, 512, 2).cuda().eval()
            feats, classes = ic(x)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert a.shape == b.shape == (6, 512)
    assert float((a - b).abs().max()) <= 2e-4 * float(a.abs().max()), float((a - b).abs().max())
    assert torch.equal(feats, b) and classes.shape == (6, 2)
    # under autograd the original (unfused) forward runs
    for p in fused.parameters():
        p.requires_grad = True
    y = fused(x[:2])
    assert y.requires_grad


@pytest.mark.parametrize("shape", [(4, 64, 112, 112), (3, 64, 56, 56), (2, 128, 28, 28), (2, 256, 14, 14), (2, 512, 7, 7),
                                   (1, 32, 3, 5), (2, 96, 1, 2), (1, 2048, 7, 7)])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (False, False)])
def test_instnorm_act_channels_last_matches_torch(shape, with_res, relu):
    """The NHWC kernel (dsmil_instnorm_act_nhwc): same operator on torch.channels_last memory, layout preserved."""
    from dsmil_wsi_b200.embedder import instnorm_act
    g = torch.Generator(device="cuda").manual_seed(sum(shape) + 1)
    x = (torch.randn(*shape, generator=g, device="cuda") * 3.0 + 40.0).contiguous(memory_format=torch.channels_last)
    res = torch.randn(*shape, generator=g, device="cuda").contiguous(memory_format=torch.channels_last) if with_res else None
    want = _ref(x.contiguous(), res.contiguous() if res is not None else None, relu)      # mean 40, std 3: the shifted sums matter
    got = instnorm_act(x, res, relu)
    assert got.is_contiguous(memory_format=torch.channels_last)
    scale = float(want.abs().max().clamp_min(1.0))
    assert float((got - want).abs().max()) <= 3e-5 * scale, shape
    x2 = x.clone(memory_format=torch.channels_last)
    out = instnorm_act(x2, res, relu, out=x2)
    assert out.data_ptr() == x2.data_ptr() and torch.equal(out, got)
    # a residual in the other layout is converted, not misread
    if res is not None:
        assert torch.equal(instnorm_act(x, res.contiguous(), relu), got)


def test_channels_last_with_odd_channel_count_still_correct():
    from dsmil_wsi_b200.embedder import instnorm_act
    x = torch.randn(2, 5, 6, 7, device="cuda").contiguous(memory_format=torch.channels_last)
    want = _ref(x.contiguous(), None, True)
    assert float((instnorm_act(x, None, True) - want).abs().max()) <= 2e-5


def test_fused_resnet18_channels_last_matches_untouched_module():
    import copy
    import torchvision.models as models
    from dsmil_wsi_b200.embedder import fuse_instance_norm
    torch.manual_seed(0)
    resnet = models.resnet18(weights=None, norm_layer=torch.nn.InstanceNorm2d)
    resnet.fc = torch.nn.Identity()
    for p in resnet.parameters():
        p.requires_grad = False
    ref = copy.deepcopy(resnet).cuda().eval()
    fused = resnet.cuda().eval().to(memory_format=torch.channels_last)
    assert fuse_instance_norm(fused) == 20
    x = torch.rand(6, 3, 224, 224, device="cuda")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            a = ref(x)
            b = fused(x.contiguous(memory_format=torch.channels_last))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert float((a - b).abs().max()) <= 2e-4 * float(a.abs().max()), float((a - b).abs().max())
