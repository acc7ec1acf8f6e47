"""Device-side losses of the reference training loop.

The reference builds the per-element BCE weights with a Python double loop that issues ~3 tiny tensor ops per
(molecule, task) pair on the host (utils.weight_tensor, utils.py:653-679: 3 072 iterations per Tox21 step) and
divides by the count of non-missing labels (train.py:326-331).  ``weighted_bce_with_logits`` computes the same
quantity with a handful of vectorised device ops and no host synchronisation; ``mse`` is train.py:321-325.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def bce_weight_table(bce_weight, n_tasks, device=None, dtype=torch.float32):
    """utils.set_weight's dict {task: [w_pos, w_neg]} (utils.py:681-700) -> tensor [n_tasks, 2]."""
    t = torch.zeros(n_tasks, 2, dtype=dtype)
    for j in range(n_tasks):
        if j in bce_weight:
            t[j, 0], t[j, 1] = float(bce_weight[j][0]), float(bce_weight[j][1])
    return t.to(device) if device is not None else t


def label_weights(table, labels):
    """utils.weight_tensor (utils.py:653-679): weight[j][0] where label == 1, weight[j][1] where label == 0,
    0 for anything else (missing labels are -1 / NaN in the reference's CSVs)."""
    lab = labels.to(table.dtype)
    lab_int = torch.nan_to_num(lab, nan=-1.0).trunc()          # int(labels[i][j]) of the reference loop
    w_pos = table[:, 0].unsqueeze(0).expand_as(lab)
    w_neg = table[:, 1].unsqueeze(0).expand_as(lab)
    zero = torch.zeros_like(lab)
    return torch.where(lab_int == 1, w_pos, torch.where(lab_int == 0, w_neg, zero))


def weighted_bce_with_logits(outputs, labels, table):
    """train.py:326-331: sum of weighted BCE-with-logits over all (molecule, task) pairs divided by the number
    of non-missing labels ((labels == 1).sum() + (labels == 0).sum())."""
    lab = labels.to(outputs.dtype)
    w = label_weights(table, lab)
    non_nan = ((lab == 1).sum() + (lab == 0).sum()).to(outputs.dtype)
    target = torch.nan_to_num(lab, nan=0.0)
    loss = F.binary_cross_entropy_with_logits(outputs.reshape(-1), target.reshape(-1), weight=w.reshape(-1),
                                              reduction="sum")
    return loss / non_nan


def mse(outputs, labels):
    """train.py:321-325."""
    return F.mse_loss(outputs.reshape(-1), labels.to(outputs.dtype).reshape(-1))
