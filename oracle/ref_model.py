"""CPU model for bench.py's reference / cpu_baseline legs.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, so the reference arm times this port: the drop-in
model's stock-torch backbone and parameter holders (plain nn.Modules, device-agnostic) with every
hot-path method rebound to the oracle's CPU restatement (oracle/recon_path.py), i.e. the same op
sequence the reference runs (model/unidefense.py:125-157, :174-255; loss/), on the host cores.
Nothing in the product package imports this file.
"""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import recon_path as O


def _act_of(m):
    return "relu" if isinstance(m, nn.ReLU) else "swish"


def _block_run(self, x, want_mean):
    mods = list(self)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.InstanceNorm2d):
            x = O.instance_norm_act(x, m.weight, m.bias, _act_of(mods[i + 1]), m.eps)
            i += 2
        elif isinstance(m, nn.Tanh):
            x = torch.tanh(x)
            i += 1
        else:
            x = m(x)
            i += 1
    return x, (x.mean(dim=(-2, -1)) if want_mean else None)


def _attention(self, pred, x, embedding):
    p = {k: v for k, v in self.named_parameters() if k.startswith(("freq_filter", "spat_filter", "fuse_coef"))}
    for k, v in self.named_buffers():
        if k.startswith(("freq_filter", "spat_filter")):
            p[k] = v
    act = _act_of(self.freq_filter.layer1[2])
    dropped = F.dropout(embedding.clone(), self.dropout.p, True) if (self.training and self.dropout.p > 0) else None
    out, fmask, smask = O.attention(pred, x, embedding, p, act, self.training, dropped, self.freq_norm)
    return {"out": out, "freq_mask": fmask, "spat_mask": smask}


def _rec_tail(self, dec_out, x, loss_dict):
    rec, spatial, freq = O.recon_tail(dec_out, x, self.freq_norm)
    loss_dict["spatial"] = spatial
    loss_dict["freq"] = freq
    return rec


def build(arch: str, **kwargs) -> nn.Module:
    """arch in {'eb4','r18','r50'} -> CPU nn.Module with the reference's forward()/loss_dict contract."""
    from unidefense_b200.model import MODEL
    from unidefense_b200.model.modules import DecoderBlock
    name = {"eb4": "UDEB4", "r18": "UDR18", "r50": "UDR50"}[arch]
    if arch == "eb4":
        kwargs.setdefault("extractor", "efficientnet-b4")
    model = MODEL[name](**kwargs)
    model.attention = types.MethodType(_attention, model)
    model._rec_tail = types.MethodType(_rec_tail, model)
    for m in model.modules():
        if isinstance(m, DecoderBlock):
            m._run = types.MethodType(_block_run, m)
    return model


def pass1_loss(out, labels, n_real, lam):
    """engine/abstract_engine.py:215-267 (first pass) with the oracle's losses."""
    ld = out["loss_dict"]
    tri = sum(O.aw_triplet_loss(f, labels) for f in ld["triplet"])
    return (F.cross_entropy(out["cls_out"], labels) + lam["mask"] * ld["freq_mask"].mean()
            + lam["mask"] * ld["spat_mask"].mean() + lam["triplet"] * tri
            + lam["recons"] * ld["spatial"][:n_real].mean() + lam["freq"] * ld["freq"][:n_real].mean())
