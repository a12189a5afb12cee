"""Drop-in UniDefense model classes: same constructor kwargs, forward() signature, return dict and
state_dict keys as the reference (model/unidefense.py:28-256 Eb4, :259-436 Res18, :439-631 Res50),
with everything downstream of the backbone features computed by the sm_100a kernels.

forward(x, pert_real_list=None, pert_fake_list=None, preserve_color=None) ->
  {'cls_out': [N,num_classes], 'rec': [N,3,R,R],
   'loss_dict': {'factorization': [N,F], 'triplet': [..[N,c]..], 'freq_mask': [N,1,h,w/2+1],
                 'spat_mask': [N,1,h,w], 'spatial': [N], 'freq': [N]}}
The backbone (EfficientNet-B4 / ResNet, SFConv included) and the bottleneck/classifier head stay
stock torch (north star).  There is no CPU path: the kernels raise on non-CUDA tensors.
"""
from functools import partial
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import perturb
from .efficientnet import EfficientNetFeatures
from .modules import (Classifier, FrequencyDynamicFilter, MemoryEfficientSwish, SpatialDynamicFilter,
                      make_decoder_block)
from .perturb import FrequencyStyleTransfer, SpatialStyleTransfer, downscale, random_blur, random_noise
from .resnet import (EmbedderRes18Layer1, EmbedderRes18Layer2, EmbedderRes50Layer1, EmbedderRes50Layer2,
                     ExtractorRes18, ExtractorRes50)

pert_noise = partial(random_noise, std=1e-4)
pert_blur = random_blur
pert_ds = downscale
PERT_FUNCS = [pert_noise, pert_blur, pert_ds]
DELIMITER_DICT = {"efficientnet-b4": [2, 6, 10, 16, 22, 30, 32]}

# decoder layouts: ('c' conv+IN+act | 't' convT s2+IN+act | 'o' conv+tanh, c_in, c_out)
def _dec_eb4():
    return [[("c", 160, 80), ("t", 80, 80), ("c", 80, 80)],
            [("c", 80, 40), ("t", 40, 40), ("c", 40, 40)],
            [("c", 40, 20), ("t", 20, 20), ("c", 20, 20), ("o", 20, 3)]]


def _dec_r18(mid):
    return [[("c", mid, 128), ("t", 128, 128), ("c", 128, 128)],
            [("c", 128, 64), ("t", 64, 64), ("c", 64, 32), ("o", 32, 3)]]


def _dec_r50(mid):
    return [[("c", mid, 256), ("t", 256, 256), ("c", 256, 256)],
            [("c", 256, 128), ("t", 128, 128), ("c", 128, 128)],
            [("c", 128, 64), ("t", 64, 64), ("c", 64, 32), ("o", 32, 3)]]


class _UniDefenseBase(nn.Module):
    """Shared hot path: augmentation dispatch, attention(), reconstruction-loss tail."""

    path = "model/unidefense.py"          # read by AbstractEngine._init_wandb (engine/abstract_engine.py:93)

    def _init_hot_path(self, att_depth, num_features, num_classes, drop_rate, activation, att_norm, affine, bias,
                       dropout_inplace):
        self.bottleneck = nn.BatchNorm1d(num_features)
        self.bottleneck.bias.requires_grad_(False)
        nn.init.constant_(self.bottleneck.weight, 1.0)
        nn.init.constant_(self.bottleneck.bias, 0.0)
        self.dropout = nn.Dropout(p=drop_rate, inplace=dropout_inplace)
        self.classifier = Classifier(num_features, num_classes)
        self.freq_filter = FrequencyDynamicFilter(att_depth, activation, att_norm, affine, bias)
        self.spat_filter = SpatialDynamicFilter(att_depth, activation, att_norm, affine, bias)
        self.freq_trans = FrequencyStyleTransfer()
        self.spat_trans = SpatialStyleTransfer()
        self.fuse_coef = nn.Parameter(torch.tensor(0.0))

    # ---- model/unidefense.py:177-200 ------------------------------------------------------
    def _perturb(self, x, pert_real_list, pert_fake_list, preserve_color):
        if not (self.training and pert_real_list is not None and pert_fake_list is not None):
            return x
        if torch.rand(1) > 0.5:
            with torch.no_grad():
                sum_real, sum_fake = len(pert_real_list), len(pert_fake_list)
                dev = x.device
                idx = torch.cat([torch.as_tensor(pert_real_list).to(dev),
                                 torch.as_tensor(pert_fake_list).to(dev) + sum_real])
                x_s = x.narrow(0, 0, sum_real + sum_fake).index_select(0, idx)
                if preserve_color:
                    x_s = ops.coral_batch(x_s, x)          # coral(s, c) for every pair, batched
                rand = torch.randint(0, 2, size=(1,))
                pert_func = self.freq_trans if rand == 0 else self.spat_trans
                return pert_func(x, x_s)
        rand = torch.randint(0, len(PERT_FUNCS), size=(1,))
        return PERT_FUNCS[int(rand)](x)

    # ---- model/unidefense.py:125-157 ------------------------------------------------------
    def attention(self, pred, x, embedding):
        size = embedding.shape[-2:]
        embedding = embedding.float()
        spat_diff, freq_diff = ops.attn_prep(pred.float(), x.float(), size, self.freq_norm)
        emb_freq = ops.rfft2_cat(embedding, self.freq_norm)
        freq_out = self.freq_filter(emb_freq, freq_diff)
        freq_filtered = ops.irfft2_cat(freq_out["out"], size, self.freq_norm)
        spat_mask = self.spat_filter.mask_only(embedding, spat_diff)
        if self.training and self.dropout.p > 0:
            residual = F.dropout(embedding.clone(), self.dropout.p, True)
        else:
            residual = None
        out = ops.attn_fuse(embedding, spat_mask, freq_filtered, residual, self.fuse_coef)
        return {"out": out, "freq_mask": freq_out["mask"], "spat_mask": spat_mask}

    # ---- model/unidefense.py:244-253 ------------------------------------------------------
    def _rec_tail(self, dec_out, x, loss_dict):
        rec, spatial, freq = ops.recon_tail(dec_out, x.float(), self.freq_norm)
        loss_dict["spatial"] = spatial
        loss_dict["freq"] = freq
        return rec


class UniDefenseModelEb4(_UniDefenseBase):
    """UniDefense model with EfficientNet backbone (model/unidefense.py:28-256)."""

    def __init__(self, extractor, extractor_weights: Optional[str] = None, bias: bool = False, drop_rate: float = 0.2,
                 affine: bool = True, num_classes: int = 1, delimiter: Optional[List] = None, freq_norm: str = "ortho",
                 **kwargs):
        super().__init__()
        self.backbone = EfficientNetFeatures(extractor, freq_norm=freq_norm, **kwargs)
        if extractor_weights is not None:
            self.backbone.load_pretrained(extractor_weights)
        self.freq_norm = freq_norm
        act, norm = MemoryEfficientSwish, nn.InstanceNorm2d
        spec = _dec_eb4()
        self.dec_block1 = make_decoder_block(spec[0], norm, act, affine, bias)
        self.dec_block2 = make_decoder_block(spec[1], norm, act, affine, bias)
        self.dec_block3 = make_decoder_block(spec[2], norm, act, affine, bias)
        self.delimiter = delimiter or DELIMITER_DICT[extractor]
        self._init_hot_path(272, self.backbone.num_features, num_classes, drop_rate, act, nn.BatchNorm2d, affine, bias,
                            dropout_inplace=True)

    def forward_backbone_block(self, x: torch.Tensor, block_id: int):
        start = self.delimiter[block_id - 1] if block_id > 0 else 0
        return self.backbone.run_blocks(x, start, self.delimiter[block_id])

    def forward(self, x, pert_real_list=None, pert_fake_list=None, preserve_color=None, **kwargs):
        loss_dict = dict()
        noise_x = self._perturb(x, pert_real_list, pert_fake_list, preserve_color)
        x_stem = self.backbone.stem(noise_x)
        x_b4 = x_stem
        for b in range(5):
            x_b4 = self.forward_backbone_block(x_b4, b)                       # [N, 160, 24, 24]
        dec_out1, tri1 = self.dec_block1.forward_with_mean(F.dropout(x_b4, 0.2, self.training))   # [N, 80, 48, 48]
        dec_out2, tri2 = self.dec_block2.forward_with_mean(dec_out1)                              # [N, 40, 96, 96]
        dec_out3 = self.dec_block3(dec_out2)                                                      # [N, 3, 192, 192]
        x_b5 = self.forward_backbone_block(x_b4, 5)                                               # [N, 272, 12, 12]
        att = self.attention(dec_out3.detach(), x, x_b5)
        x_out = self.forward_backbone_block(att["out"], 6)
        x_out = self.backbone._avg_pooling(self.backbone.head(x_out)).flatten(1)
        x_out = self.bottleneck(x_out.float())
        loss_dict["factorization"] = x_out
        x_out = self.dropout(x_out)            # inplace: 'factorization' aliases the dropped tensor (Appendix D)
        loss_dict["triplet"] = [x_b4.float().mean([-2, -1]), tri1, tri2]
        loss_dict["freq_mask"] = att["freq_mask"]
        loss_dict["spat_mask"] = att["spat_mask"]
        cls_out = self.classifier(x_out)
        rec = self._rec_tail(dec_out3, x, loss_dict)
        return {"cls_out": cls_out, "rec": rec, "loss_dict": loss_dict}


class UniDefenseModelRes18(_UniDefenseBase):
    """UniDefense model with ResNet18 backbone (model/unidefense.py:259-436)."""

    def __init__(self, extractor="resnet18", extractor_weights: Optional[str] = None, mid_depth=448, bias: bool = False,
                 drop_rate: float = 0.2, affine: bool = True, num_classes: int = 2, freq_norm: str = "ortho", **kwargs):
        super().__init__()
        enc_norm, act = nn.BatchNorm2d, nn.ReLU
        self.freq_norm = freq_norm
        self.extractor = ExtractorRes18(extractor, extractor_weights, freq_norm)
        self.emb_block1 = EmbedderRes18Layer1(mid_depth, bias, enc_norm, affine, act)
        self.emb_block2 = EmbedderRes18Layer2(bias, enc_norm, affine, act)
        spec = _dec_r18(mid_depth)
        self.dec_block1 = make_decoder_block(spec[0], nn.InstanceNorm2d, act, affine, bias)
        self.dec_block2 = make_decoder_block(spec[1], nn.InstanceNorm2d, act, affine, bias)
        self._init_hot_path(512, 512, num_classes, drop_rate, act, enc_norm, affine, bias, dropout_inplace=False)

    def forward(self, x, pert_real_list=None, pert_fake_list=None, preserve_color=None, **kwargs):
        loss_dict = dict()
        noise_x = self._perturb(x, pert_real_list, pert_fake_list, preserve_color)
        _, ext_feat = self.extractor(noise_x)                                                     # [N, 448, R/8, R/8]
        dec_out1, tri1 = self.dec_block1.forward_with_mean(F.dropout(ext_feat, 0.2, self.training))
        dec_out2 = self.dec_block2(dec_out1)                                                      # [N, 3, R/2, R/2]
        emb_feat = self.emb_block1(ext_feat)                                                      # [N, 512, R/16, R/16]
        att = self.attention(dec_out2.detach(), x, emb_feat)
        emb_feat = self.emb_block2(att["out"])
        emb_feat = F.adaptive_avg_pool2d(emb_feat, 1).flatten(1)
        emb_feat = self.bottleneck(emb_feat.float())
        loss_dict["factorization"] = emb_feat
        emb_feat = self.dropout(emb_feat)
        loss_dict["triplet"] = [ext_feat.float().mean([-2, -1]), tri1]
        loss_dict["freq_mask"] = att["freq_mask"]
        loss_dict["spat_mask"] = att["spat_mask"]
        cls_out = self.classifier(emb_feat)
        rec = self._rec_tail(dec_out2, x, loss_dict)
        return {"cls_out": cls_out, "rec": rec, "loss_dict": loss_dict}


class UniDefenseModelRes50(_UniDefenseBase):
    """UniDefense model with ResNet50 backbone (model/unidefense.py:439-631)."""

    def __init__(self, extractor="resnet50", extractor_weights: Optional[str] = None, mid_depth=1024, bias: bool = False,
                 drop_rate: float = 0.2, affine: bool = True, num_classes: int = 2, freq_norm: str = "ortho", **kwargs):
        super().__init__()
        enc_norm, act = nn.BatchNorm2d, nn.ReLU
        self.freq_norm = freq_norm
        self.extractor = ExtractorRes50(extractor, extractor_weights, freq_norm)
        self.emb_block1 = EmbedderRes50Layer1(mid_depth, bias, enc_norm, affine, act)
        self.emb_block2 = EmbedderRes50Layer2(bias, enc_norm, affine, act)
        spec = _dec_r50(mid_depth)
        self.dec_block1 = make_decoder_block(spec[0], nn.InstanceNorm2d, act, affine, bias)
        self.dec_block2 = make_decoder_block(spec[1], nn.InstanceNorm2d, act, affine, bias)
        self.dec_block3 = make_decoder_block(spec[2], nn.InstanceNorm2d, act, affine, bias)
        self._init_hot_path(2048, 2048, num_classes, drop_rate, act, enc_norm, affine, bias, dropout_inplace=False)

    def forward(self, x, pert_real_list=None, pert_fake_list=None, preserve_color=None, **kwargs):
        loss_dict = dict()
        noise_x = self._perturb(x, pert_real_list, pert_fake_list, preserve_color)
        ext_feat = self.extractor(noise_x)                                                        # [N, 1024, R/16, R/16]
        dec_out1, tri1 = self.dec_block1.forward_with_mean(F.dropout(ext_feat, 0.2, self.training))
        dec_out2 = self.dec_block2(dec_out1)
        dec_out3 = self.dec_block3(dec_out2)                                                      # [N, 3, R/2, R/2]
        emb_feat = self.emb_block1(ext_feat)                                                      # [N, 2048, R/32, R/32]
        att = self.attention(dec_out3.detach(), x, emb_feat)
        emb_feat = self.emb_block2(att["out"])
        emb_feat = F.adaptive_avg_pool2d(emb_feat, 1).flatten(1)
        emb_feat = self.bottleneck(emb_feat.float())
        loss_dict["factorization"] = emb_feat
        emb_feat = self.dropout(emb_feat)
        loss_dict["triplet"] = [ext_feat.float().mean([-2, -1]), tri1]
        loss_dict["freq_mask"] = att["freq_mask"]
        loss_dict["spat_mask"] = att["spat_mask"]
        cls_out = self.classifier(emb_feat)
        rec = self._rec_tail(dec_out3, x, loss_dict)
        return {"cls_out": cls_out, "rec": rec, "loss_dict": loss_dict}
