"""EfficientNet feature extractor with SFConv depthwise stages (stock torch backbone).

Mirrors the module tree / state_dict names of the reference's lukemelas-EfficientNet fork
(model/efficientnet/model.py:36-143 MBConvBlock, :146-260 EfficientNet) so reference
checkpoints load strictly: `_conv_stem`, `_bn0`, `_blocks.{i}.{_expand_conv,_bn0,_depthwise_conv,
_bn1,_se_reduce,_se_expand,_project_conv,_bn2}`, `_conv_head`, `_bn1`.  Written from the
published architecture (Tan & Le 2019: compound scaling table, MBConv, SE ratio 0.25, BN
momentum 0.99 / eps 1e-3, stochastic depth); SFConv replaces the depthwise conv in every stage
except the first two and the last (model/efficientnet/model.py:205,214).
"""
import math
from dataclasses import dataclass, replace
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sfconv import SamePadConv2d, SFSamePadConv2d

# name -> (width, depth, resolution, dropout)
SCALING = {
    "efficientnet-b0": (1.0, 1.0, 224, 0.2), "efficientnet-b1": (1.0, 1.1, 240, 0.2),
    "efficientnet-b2": (1.1, 1.2, 260, 0.3), "efficientnet-b3": (1.2, 1.4, 300, 0.3),
    "efficientnet-b4": (1.4, 1.8, 380, 0.4), "efficientnet-b5": (1.6, 2.2, 456, 0.4),
    "efficientnet-b6": (1.8, 2.6, 528, 0.5), "efficientnet-b7": (2.0, 3.1, 600, 0.5),
}


@dataclass(frozen=True)
class Stage:
    repeat: int
    kernel: int
    stride: int
    expand: int
    cin: int
    cout: int
    se: float = 0.25


BASE_STAGES = [Stage(1, 3, 1, 1, 32, 16), Stage(2, 3, 2, 6, 16, 24), Stage(2, 5, 2, 6, 24, 40), Stage(3, 3, 2, 6, 40, 80),
               Stage(3, 5, 1, 6, 80, 112), Stage(4, 5, 2, 6, 112, 192), Stage(1, 3, 1, 6, 192, 320)]
BN_MOMENTUM, BN_EPS = 0.01, 1e-3


def scale_width(ch: int, width: float, divisor: int = 8) -> int:
    ch = ch * width
    new = max(divisor, int(ch + divisor / 2) // divisor * divisor)
    if new < 0.9 * ch:
        new += divisor
    return int(new)


def scale_depth(repeat: int, depth: float) -> int:
    return int(math.ceil(depth * repeat))


def stochastic_depth(x, p: float, training: bool):
    """Per-sample residual-branch drop ('drop connect', model/efficientnet/utils.py:131-151)."""
    if not training or not p:
        return x
    keep = 1.0 - p
    gate = torch.floor(keep + torch.rand([x.shape[0], 1, 1, 1], dtype=x.dtype, device=x.device))
    return x / keep * gate


class MBConvBlock(nn.Module):
    def __init__(self, st: Stage, image_size: int, freq_norm: Optional[str]):
        super().__init__()
        self.stride, self.cin, self.cout, self.expand = st.stride, st.cin, st.cout, st.expand
        mid = st.cin * st.expand
        if st.expand != 1:
            self._expand_conv = SamePadConv2d(st.cin, mid, 1, image_size=image_size, bias=False)
            self._bn0 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM, eps=BN_EPS)
        if freq_norm is not None:
            self._depthwise_conv = SFSamePadConv2d(mid, mid, st.kernel, st.stride, image_size=image_size,
                                                   freq_norm=freq_norm, groups=mid, bias=False)
        else:
            self._depthwise_conv = SamePadConv2d(mid, mid, st.kernel, st.stride, image_size=image_size, groups=mid,
                                                 bias=False)
        self._bn1 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM, eps=BN_EPS)
        out_size = int(math.ceil(image_size / st.stride))
        self.has_se = st.se is not None and 0 < st.se <= 1
        if self.has_se:
            squeezed = max(1, int(st.cin * st.se))
            self._se_reduce = SamePadConv2d(mid, squeezed, 1, image_size=1)
            self._se_expand = SamePadConv2d(squeezed, mid, 1, image_size=1)
        self._project_conv = SamePadConv2d(mid, st.cout, 1, image_size=out_size, bias=False)
        self._bn2 = nn.BatchNorm2d(st.cout, momentum=BN_MOMENTUM, eps=BN_EPS)

    def forward(self, inputs, drop_connect_rate=None):
        x = inputs
        if self.expand != 1:
            x = F.silu(self._bn0(self._expand_conv(x)))
        x = F.silu(self._bn1(self._depthwise_conv(x)))
        if self.has_se:
            s = F.adaptive_avg_pool2d(x, 1)
            s = self._se_expand(F.silu(self._se_reduce(s)))
            x = torch.sigmoid(s) * x
        x = self._bn2(self._project_conv(x))
        if self.stride == 1 and self.cin == self.cout:
            x = stochastic_depth(x, drop_connect_rate, self.training) + inputs
        return x


class EfficientNetFeatures(nn.Module):
    """Stem + MBConv blocks + head conv/BN (no classifier: include_top=False in the reference call,
    model/unidefense.py:45-52)."""

    def __init__(self, name: str = "efficientnet-b4", freq_norm: Optional[str] = "ortho", drop_connect_rate: float = 0.2,
                 image_size: Optional[int] = None):
        super().__init__()
        if name not in SCALING:
            raise ValueError("model_name should be one of: " + ", ".join(SCALING))
        width, depth, res, _ = SCALING[name]
        size = image_size or res
        self.drop_connect_rate = drop_connect_rate
        stem = scale_width(32, width)
        self._conv_stem = SamePadConv2d(3, stem, 3, 2, image_size=size, bias=False)
        self._bn0 = nn.BatchNorm2d(stem, momentum=BN_MOMENTUM, eps=BN_EPS)
        size = int(math.ceil(size / 2))
        blocks: List[nn.Module] = []
        self.stage_ends: List[int] = []
        last = len(BASE_STAGES) - 1
        for si, base in enumerate(BASE_STAGES):
            st = replace(base, cin=scale_width(base.cin, width), cout=scale_width(base.cout, width),
                         repeat=scale_depth(base.repeat, depth))
            fn = freq_norm if si not in (0, 1, last) else None
            blocks.append(MBConvBlock(st, size, fn))
            size = int(math.ceil(size / st.stride))
            for _ in range(st.repeat - 1):
                blocks.append(MBConvBlock(replace(st, cin=st.cout, stride=1), size, fn))
            self.stage_ends.append(len(blocks))
        self._blocks = nn.ModuleList(blocks)
        head = scale_width(1280, width)
        self._conv_head = SamePadConv2d(blocks[-1].cout, head, 1, image_size=size, bias=False)
        self._bn1 = nn.BatchNorm2d(head, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self.num_features = head

    def load_pretrained(self, weights_path: str):
        """lukemelas-format checkpoint (model/efficientnet/utils.py:589-623): the classifier keys are
        dropped, only SFConv-owned keys (freq_conv / sf_coef) may be missing."""
        sd = torch.load(weights_path, map_location="cpu")
        sd.pop("_fc.weight", None)
        sd.pop("_fc.bias", None)
        ret = self.load_state_dict(sd, strict=False)
        bad = [k for k in ret.missing_keys if "sf_coef" not in k and "freq_conv" not in k]
        if bad:
            raise RuntimeError("Missing keys when loading pretrained weights: {}".format(bad))
        assert not ret.unexpected_keys, "Missing keys when loading pretrained weights: {}".format(ret.unexpected_keys)
        print("Loaded pretrained weights for efficientnet from {}".format(weights_path))

    def stem(self, x):
        return F.silu(self._bn0(self._conv_stem(x)))

    def run_blocks(self, x, start: int, end: int):
        n = len(self._blocks)
        for i in range(start, end):
            rate = self.drop_connect_rate * float(i) / n if self.drop_connect_rate else None
            x = self._blocks[i](x, drop_connect_rate=rate)
        return x

    def head(self, x):
        return F.silu(self._bn1(self._conv_head(x)))
