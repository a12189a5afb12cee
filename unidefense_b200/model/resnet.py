"""ResNet-18/50 extractors and embedder blocks with SFConv (stock torch backbone).

Mirrors the module tree / state_dict names of the reference's timm-ResNet fork
(model/resnet/exp.py:79-236 BasicBlock/Bottleneck, :284-327 make_blocks, :437-452 ResNet ctor)
and of model/resnet/module_exp.py:8-177 (ExtractorRes18/50, EmbedderRes{18,50}Layer{1,2}), so
torchvision/timm ResNet checkpoints and reference checkpoints load strictly.  Written from the
published architecture (He et al. 2015, v1.5 stride placement).  SFConv rule of the reference:
inside stages 2-4, a conv becomes an SFConv2d when its input and output widths are equal
(exp.py:95-98,:107-110,:167-190, stage gate :303); embedder-block SFConvs get freq_norm=None
(module_exp.py:68,94,120,157) -> un-normalised forward FFT, 1/(HW) inverse.
"""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sfconv import SFConv2d


def _conv(cin, cout, k, stride, padding, freq_norm):
    if freq_norm is not None and cin == cout:
        return SFConv2d(cin, cout, k, stride, freq_norm=freq_norm, padding=padding, bias=False)
    return nn.Conv2d(cin, cout, k, stride, padding, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, freq_norm=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride, 1, freq_norm)
        self.bn1 = nn.BatchNorm2d(planes)
        self.drop_block = nn.Identity()
        self.act1 = nn.ReLU(inplace=True)
        self.aa = nn.Identity()
        self.conv2 = _conv(planes, planes, 3, 1, 1, freq_norm)
        self.bn2 = nn.BatchNorm2d(planes)
        self.act2 = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.act2(y + shortcut)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, freq_norm=None):
        super().__init__()
        out = planes * self.expansion
        self.conv1 = _conv(inplanes, planes, 1, 1, 0, freq_norm)
        self.bn1 = nn.BatchNorm2d(planes)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = _conv(planes, planes, 3, stride, 1, freq_norm)
        self.bn2 = nn.BatchNorm2d(planes)
        self.drop_block = nn.Identity()
        self.act2 = nn.ReLU(inplace=True)
        self.aa = nn.Identity()
        self.conv3 = _conv(planes, out, 1, 1, 0, freq_norm)
        self.bn3 = nn.BatchNorm2d(out)
        self.act3 = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.bn1(self.conv1(x)))
        y = self.act2(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.act3(y + shortcut)


def _make_stage(block, inplanes, planes, blocks, stride, freq_norm):
    down = None
    if stride != 1 or inplanes != planes * block.expansion:
        down = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, 1, stride, 0, bias=False),
                             nn.BatchNorm2d(planes * block.expansion))
    layers = [block(inplanes, planes, stride, down, freq_norm)]
    for _ in range(1, blocks):
        layers.append(block(planes * block.expansion, planes, 1, None, freq_norm))
    return nn.Sequential(*layers)


def _init_resnet(module, zero_last=True):
    """kaiming fan_out for every conv (SFConv's freq_conv included, it is an nn.Conv2d), BN to (1, 0), last BN of each
    residual block zeroed (exp.py:454-464)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
    if zero_last:
        for m in module.modules():
            if isinstance(m, BasicBlock):
                nn.init.zeros_(m.bn2.weight)
            elif isinstance(m, Bottleneck):
                nn.init.zeros_(m.bn3.weight)


def _load_backbone_weights(net, weights_path, what):
    """custom_resnet18/50 (exp.py:505-535): strict except for the SFConv-only keys and the unused head."""
    sd = torch.load(weights_path, map_location="cpu")
    own = net.state_dict()
    unexpected = [k for k in sd if k not in own and not k.startswith(("layer4.", "fc."))]
    if unexpected:
        raise RuntimeError(f"Unexpected keys when loading pretrained weights: {unexpected}")
    missing = [k for k in own if k not in sd and "sf_coef" not in k and "freq_conv" not in k]
    if missing:
        raise RuntimeError(f"Missing keys when loading pretrained weights: {missing}")
    net.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    print(f"Loaded pretrained weights for {what} from {weights_path}.")


class _Extractor(nn.Module):
    def __init__(self, block, layers, pretrained, freq_norm):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = _make_stage(block, 64, 64, layers[0], 1, None)
        self.layer2 = _make_stage(block, 64 * block.expansion, 128, layers[1], 2, freq_norm)
        self.layer3 = _make_stage(block, 128 * block.expansion, 256, layers[2], 2, freq_norm)
        _init_resnet(self)
        if pretrained is not None:
            _load_backbone_weights(self, pretrained, type(self).__name__)


class ExtractorRes18(_Extractor):
    """module_exp.py:8-32: no max-pool; returns (layer3, cat[pool(layer1), pool(layer2), layer3]) -> 448 channels."""

    def __init__(self, extractor="resnet18", pretrained=None, freq_norm=None):
        if extractor != "resnet18":
            raise ValueError(f"ExtractorRes18 supports 'resnet18', got {extractor!r}")
        super().__init__(BasicBlock, [2, 2, 2], pretrained, freq_norm)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        p1 = self.layer1(x)
        p2 = self.layer2(p1)
        p3 = self.layer3(p2)
        size = p3.shape[-2:]
        return p3, torch.cat([F.adaptive_avg_pool2d(p1, size), F.adaptive_avg_pool2d(p2, size), p3], dim=1)


class ExtractorRes50(_Extractor):
    """module_exp.py:35-59: standard stem with max-pool, returns layer3 (1024 channels)."""

    def __init__(self, extractor="resnet50", pretrained=None, freq_norm=None):
        if extractor != "resnet50":
            raise ValueError(f"ExtractorRes50 supports 'resnet50', got {extractor!r}")
        super().__init__(Bottleneck, [3, 4, 6], pretrained, freq_norm)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer3(self.layer2(self.layer1(x)))


def _down(in_depth, out_depth, bias, norm, affine):
    return nn.Sequential(nn.Conv2d(in_depth, out_depth, 1, bias=bias), norm(out_depth, affine=affine),
                         nn.MaxPool2d(kernel_size=3, stride=2, padding=1))


class EmbedderRes18Layer1(nn.Module):
    """module_exp.py:62-87."""

    def __init__(self, in_depth, bias, norm, affine, activation):
        super().__init__()
        self.conv1 = nn.Conv2d(in_depth, 512, 3, 2, padding=1, bias=bias)
        self.norm1 = norm(512, affine=affine)
        self.act = activation(inplace=True)
        self.conv2 = SFConv2d(512, 512, 3, 1, padding=1, bias=bias)
        self.norm2 = norm(512, affine=affine)
        self.downsample = _down(in_depth, 512, bias, norm, affine)

    def forward(self, x):
        y = self.act(self.norm1(self.conv1(x)))
        y = self.norm2(self.conv2(y))
        return self.act(y + self.downsample(x))


class EmbedderRes18Layer2(nn.Module):
    """module_exp.py:90-111."""

    def __init__(self, bias, norm, affine, activation):
        super().__init__()
        self.conv1 = SFConv2d(512, 512, 3, 1, padding=1, bias=bias)
        self.norm1 = norm(512, affine=affine)
        self.act = activation(inplace=True)
        self.conv2 = nn.Conv2d(512, 512, 3, 1, padding=1, bias=bias)
        self.norm2 = norm(512, affine=affine)

    def forward(self, x):
        y = self.act(self.norm1(self.conv1(x)))
        y = self.norm2(self.conv2(y))
        return self.act(y + x)


class EmbedderRes50Layer1(nn.Module):
    """module_exp.py:114-148."""

    def __init__(self, in_depth, bias, norm, affine, activation):
        super().__init__()
        self.conv1 = nn.Conv2d(in_depth, 512, kernel_size=1, bias=bias)
        self.norm1 = norm(512, affine=affine)
        self.act = activation(inplace=True)
        self.conv2 = SFConv2d(512, 512, 3, 2, padding=1, bias=bias)
        self.norm2 = norm(512, affine=affine)
        self.conv3 = nn.Conv2d(512, 2048, kernel_size=1, bias=bias)
        self.norm3 = norm(2048, affine=affine)
        self.downsample = _down(in_depth, 2048, bias, norm, affine)

    def forward(self, x):
        y = self.act(self.norm1(self.conv1(x)))
        y = self.act(self.norm2(self.conv2(y)))
        y = self.norm3(self.conv3(y))
        return self.act(y + self.downsample(x))


class EmbedderRes50Layer2(nn.Module):
    """module_exp.py:151-177."""

    def __init__(self, bias, norm, affine, activation):
        super().__init__()
        self.conv1 = nn.Conv2d(2048, 512, kernel_size=1, bias=bias)
        self.norm1 = norm(512, affine=affine)
        self.act = activation(inplace=True)
        self.conv2 = SFConv2d(512, 512, 3, stride=1, padding=1, bias=bias)
        self.norm2 = norm(512, affine=affine)
        self.conv3 = nn.Conv2d(512, 2048, kernel_size=1, bias=bias)
        self.norm3 = norm(2048, affine=affine)

    def forward(self, x):
        y = self.act(self.norm1(self.conv1(x)))
        y = self.act(self.norm2(self.conv2(y)))
        y = self.norm3(self.conv3(y))
        return self.act(y + x)
