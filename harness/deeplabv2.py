"""Test / bench infrastructure, NOT the product: a random-init DeepLabv2-ResNet101 two-head network that stands in
for the reference's ``DeeplabMulti`` (``graphs/models/deeplab_multi.py:69-187``) as the PRODUCER of the hot path's
inputs in the config-5 adaptation step (``tools/solve_crosscity.py:165-249``).  The backbone is out of scope (it stays
on cuDNN, north star); this file exists because the reference's model cannot travel to the GPU box.

Same topology and the same parameter names as the reference (its ``state_dict`` loads here unchanged --
``tests/test_harness_model.py`` checks outputs against the reference model where ``/root/reference`` is present):
7x7/2 stem, ceil-mode 3x3/2 max-pool, bottleneck stages [3, 4, 23, 3] at strides (1, 2, 1, 1) and dilations (1, 1, 2, 4),
``layer5`` on stage 3 and ``layer6`` on stage 4, each four dilated 3x3 classifiers (6/12/18/24) of which -- like the
reference, whose ``Classifier_Module.forward`` returns inside its loop (``deeplab_multi.py:62-66``) -- only the first two
are summed.  The one deliberate difference: ``forward(x, upsample=False)`` can hand back the LOW-resolution head logits,
which is what the fused kernels consume; ``upsample=True`` reproduces the reference's two ``F.interpolate`` calls
(``deeplab_multi.py:124,128``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

STAGES = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 23, 1, 2), (512, 3, 1, 4))       # planes, blocks, stride, dilation
ASPP_RATES = (6, 12, 18, 24)


def head_size(H, W):
    """(h, w) of the two heads for an H x W input: stem conv /2, ceil-mode pool /2, stage-2 stride 2."""
    def one(n):
        n = (n - 1) // 2 + 1                  # 7x7 stride 2 pad 3
        n = -(-(n - 1) // 2) + 1              # 3x3 stride 2 pad 1, ceil_mode
        return (n - 1) // 2 + 1               # 1x1 stride 2
    return one(H), one(W)


class _Block(nn.Module):
    def __init__(self, cin, planes, stride, dilation, project):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, 4 * planes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(4 * planes)
        self.downsample = None
        if project:
            self.downsample = nn.Sequential(nn.Conv2d(cin, 4 * planes, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(4 * planes))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Head(nn.Module):
    def __init__(self, cin, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList(nn.Conv2d(cin, num_classes, 3, padding=r, dilation=r) for r in ASPP_RATES)

    def forward(self, x):
        return self.conv2d_list[0](x) + self.conv2d_list[1](x)          # branches 2 and 3 exist but are never used


class DeepLabV2Harness(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        for p in self.bn1.parameters():
            p.requires_grad = False
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1, ceil_mode=True)
        cin = 64
        for i, (planes, blocks, stride, dil) in enumerate(STAGES, start=1):
            mods = [_Block(cin, planes, stride, dil, True)]
            cin = 4 * planes
            mods += [_Block(cin, planes, 1, dil, False) for _ in range(blocks - 1)]
            setattr(self, f"layer{i}", nn.Sequential(*mods))
        self.layer5 = _Head(1024, num_classes)
        self.layer6 = _Head(2048, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0.0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x, upsample=False):
        size = x.shape[2:]
        x = self.maxpool(F.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(self.layer1(x)))
        aux_head = self.layer5(x)
        main_head = self.layer6(self.layer4(x))
        if upsample:
            main_head, aux_head = (F.interpolate(t, size=size, mode="bilinear", align_corners=True)
                                   for t in (main_head, aux_head))
        return main_head, aux_head            # (pred, pred_2) in the trainers' naming
