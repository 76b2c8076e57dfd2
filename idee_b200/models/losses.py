"""Loss modules of the synthetic training step: drop-ins for ``models/losses.py:98-168`` backed by fused CUDA kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class BCE_loss_synthetic(nn.Module):
    """Class-frequency weighted BCE-with-logits (losses.py:98-124)."""

    def forward(self, pred, target):
        """pred, target [N,C,H,W] -> scalar"""
        return ops.bce_loss_map(pred, target)


class Anomaly_L1_loss_synthetic(nn.Module):
    """Masked L1 between z_q and the code of index 0 (losses.py:127-168)."""

    def __init__(self, n_dynamic: int = 3, delta_t: int = 8, dim: int = 24):
        super().__init__()
        self.n_dynamic, self.delta_t, self.dim = n_dynamic, delta_t, dim

    def forward(self, pred, mask_extreme, vq_0):
        """pred z_q [N,V,C,T,H,W], mask_extreme [N,H,W], vq_0 [1,C] -> scalar"""
        rank1 = getattr(pred, "_idee_rank1", None)
        if rank1 is not None and pred.shape[2] == 16:
            # z_q as returned by idee_b200's VQ_model carries its rank-1 factors (x, w_out, b_out): same loss, 1/16 of the bytes
            xq, w_out, b_out = rank1
            return ops.AnomalyRank1.apply(xq, w_out, b_out, mask_extreme, vq_0.reshape(-1))
        tok = pred.permute(0, 1, 3, 4, 5, 2)                          # [N,V,T,H,W,C]; a view when pred is channel-last
        return ops.AnomalyL1.apply(tok, mask_extreme, vq_0.reshape(-1))


def train_step_loss(model, data_d, mask_extreme, mask_extreme_loss, lambda_anomaly: float = 100.0):
    """Loss assembly of one optimisation step exactly as train_synthetic.py:175-201; returns (total, outputs)."""
    crit, crit_an = BCE_loss_synthetic(), Anomaly_L1_loss_synthetic()
    pred, pred_y, anomaly, z_q, loss_z_q = model(data_d)
    tgt = mask_extreme.unsqueeze(1).float()
    loss = crit(pred, tgt)
    vq = model.module.vq if hasattr(model, "module") else model.vq
    vq0 = vq.indices_to_codes(torch.zeros(1, dtype=torch.long, device=data_d.device)).clone().detach()
    loss_anomaly = crit_an(z_q, mask_extreme_loss.float(), vq0)
    stacked = getattr(pred_y, "stacked", None)
    if stacked is not None and stacked.dtype == torch.float32:
        # the V per-variable maps are views of one [N,V,H,W] tensor: same V losses (train_synthetic.py:196-198), one launch set
        loss_var = ops.bce_loss_maps(stacked, tgt).sum()
    else:
        loss_var = 0
        for y in pred_y:
            loss_var = loss_var + crit(y, tgt)
    total = loss + loss_anomaly * lambda_anomaly + loss_var + loss_z_q
    return total, dict(pred=pred, pred_y=pred_y, anomaly=anomaly, z_q=z_q, loss_z_q=loss_z_q, loss_bce=loss,
                       loss_anomaly=loss_anomaly, loss_var=loss_var)
