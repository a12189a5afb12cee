"""Loss registry with the reference's names and call signatures (loss/__init__.py:6-18), backed
by the fused sm_100a kernels (forward + gradient in one launch, no host syncs)."""
import torch
import torch.nn as nn

from .. import ops


class AsymmetricalWeightedTripletLoss(nn.Module):
    """loss/triplet_loss.py:69-82.  labels: int64, label-0 (real) rows first."""

    def forward(self, global_feat, labels, normalize_feature=False):
        if normalize_feature:
            global_feat = global_feat / (torch.norm(global_feat, 2, -1, keepdim=True) + 1e-12)
        return ops.triplet_loss(global_feat.float(), labels)


class FactorizationLoss(nn.Module):
    """loss/calib_loss.py:5-28."""

    def __init__(self, off_diag_weight=0.005):
        super().__init__()
        self.off_diag_weight = off_diag_weight

    def forward(self, emb_a, emb_b, eps=1e-6):
        return ops.factorization_loss(emb_a.float(), emb_b.float(), self.off_diag_weight, eps)


class KLDivLogTarget(nn.Module):
    """nn.KLDivLoss(reduction='batchmean', log_target=True) as the engine calls it
    (engine/abstract_engine.py:337,345): both arguments are log-probabilities [N, M]."""

    def forward(self, log_pred, log_target):
        return ops.kl_div_log_target(log_pred.float(), log_target.float())


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() (loss/__init__.py:14) for the engine's call `softmax(cls_out, in_tgt)`
    (engine/abstract_engine.py:259): class-index targets, mean reduction; soft targets / weights / label smoothing are
    not on the reference path and go to torch."""

    def forward(self, logits, target):
        if logits.is_cuda and logits.dim() == 2 and target.dtype == torch.int64 and target.dim() == 1:
            return ops.cross_entropy(logits, target)
        return nn.functional.cross_entropy(logits, target)


class BCEWithLogitsLoss(nn.Module):
    """nn.BCEWithLogitsLoss() (loss/__init__.py:12), the num_classes == 1 branch (engine/abstract_engine.py:257)."""

    def forward(self, logits, target):
        if logits.is_cuda and logits.shape == target.shape:
            return ops.bce_with_logits(logits, target)
        return nn.functional.binary_cross_entropy_with_logits(logits, target)


LOSSES = {
    "mse": nn.MSELoss(),
    "bce": BCEWithLogitsLoss(),
    "factorization": FactorizationLoss(),
    "cross_entropy": CrossEntropyLoss(),
    "aw_triplet": AsymmetricalWeightedTripletLoss(),
    "kl_div": KLDivLogTarget(),
}


def get_loss(name="cross_entropy", device="cuda:0"):
    print(f"Using loss: '{LOSSES[name]}'")
    return LOSSES[name].to(device)
