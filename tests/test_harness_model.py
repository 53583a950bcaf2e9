"""The config-5 harness network (harness/deeplabv2.py, test infrastructure) against the reference's DeeplabMulti where
the reference tree is present (the build container); on the GPU box only the shape checks run."""
import importlib.util
import os

import pytest
import torch
import torch.nn.functional as F

from harness.deeplabv2 import DeepLabV2Harness, head_size

REF = "/root/reference/graphs/models/deeplab_multi.py"


def test_head_sizes_of_the_baseline_configs():
    # verified there by running the reference model
    assert head_size(512, 1024) == (65, 129)
    assert head_size(720, 1280) == (91, 161)
    assert head_size(760, 1280) == (96, 161)
    assert head_size(640, 1280) == (81, 161)


def test_lowres_heads_and_upsample_mode_agree():
    torch.manual_seed(0)
    net = DeepLabV2Harness(13).eval()
    x = torch.randn(1, 3, 65, 97)
    with torch.no_grad():
        lo1, lo2 = net(x)
        up1, up2 = net(x, upsample=True)
    assert tuple(lo1.shape[2:]) == head_size(65, 97) == tuple(lo2.shape[2:])
    for lo, up in ((lo1, up1), (lo2, up2)):
        assert torch.equal(F.interpolate(lo, size=(65, 97), mode="bilinear", align_corners=True), up)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_same_function_as_the_reference_model():
    spec = importlib.util.spec_from_file_location("ref_deeplab_multi", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(12345)
    ref = mod.DeeplabMulti(num_classes=13, pretrained=False)
    mine = DeepLabV2Harness(13)
    assert sorted(mine.state_dict().keys()) == sorted(ref.state_dict().keys())
    assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in ref.parameters())
    assert [n for n, p in mine.named_parameters() if not p.requires_grad] == \
           [n for n, p in ref.named_parameters() if not p.requires_grad]
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(1, 3, 65, 129)
    for train in (False, True):            # eval and train mode (batch statistics) alike
        ref.train(train), mine.train(train)
        with torch.no_grad():
            r2, r1 = ref(x)
            m2, m1 = mine(x, upsample=True)
        assert torch.allclose(m2, r2, rtol=1e-5, atol=1e-7) and torch.allclose(m1, r1, rtol=1e-5, atol=1e-7)
