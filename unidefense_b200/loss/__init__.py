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


LOSSES = {
    "mse": nn.MSELoss(),
    "bce": nn.BCEWithLogitsLoss(),
    "factorization": FactorizationLoss(),
    "cross_entropy": nn.CrossEntropyLoss(),
    "aw_triplet": AsymmetricalWeightedTripletLoss(),
    "kl_div": KLDivLogTarget(),
}


def get_loss(name="cross_entropy", device="cuda:0"):
    print(f"Using loss: '{LOSSES[name]}'")
    return LOSSES[name].to(device)
