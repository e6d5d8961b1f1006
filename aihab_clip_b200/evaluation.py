"""Scoring metrics consumed after the hot path — drop-in for the torch-only parts of
``aihab_utils/evaluation.py`` and ``methods/utils.py:16-21``.  Same names, arguments, return values and error
behaviour; plotting / W&B reporting (draw_cm, save_classification) is out of scope.  Inputs are plain
``torch.Tensor`` logits ``[N, C]`` and int64 targets on any device (the logits the scoring kernel produces)."""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F


def _fused_ok(logits: torch.Tensor, num_l2: int) -> bool:
    """fp32 CUDA logits within the limits of the fused metrics kernel (aihab_l2_metrics); other inputs take the
    reference's torch formulation below, which also defines the semantics for fp16 / CPU tensors."""
    return (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and 1 <= logits.shape[1] <= 1024
            and 1 <= num_l2 <= 256)


def cls_acc(output: torch.Tensor, target: torch.Tensor, topk: int = 1) -> float:
    """ref methods/utils.py:16-21 — top-k accuracy in percent."""
    pred = output.topk(topk, 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return 100.0 * float(correct[:topk].reshape(-1).float().sum().item()) / target.shape[0]


def map_l3_targets_to_l2(targets_l3: torch.Tensor, l3_to_l2: Union[Sequence[int], torch.Tensor]) -> torch.Tensor:
    """ref aihab_utils/evaluation.py:80-89."""
    if torch.is_tensor(l3_to_l2):
        lut = l3_to_l2.to(device=targets_l3.device, dtype=torch.long)
    else:
        lut = torch.tensor(list(l3_to_l2), device=targets_l3.device, dtype=torch.long)
    return lut[targets_l3.long()]


def aggregate_logits_to_l2(logits_l3: torch.Tensor, l3_to_l2: Union[Sequence[int], torch.Tensor], num_l2: int,
                           reduce: str = "mean") -> torch.Tensor:
    """ref aihab_utils/evaluation.py:92-142 — sum / mean / logsumexp of L3 logits per L2 group, accumulated in L3
    id order like the reference (so CPU results are bit-identical)."""
    l3_list = l3_to_l2.detach().cpu().tolist() if torch.is_tensor(l3_to_l2) else list(l3_to_l2)
    if int(logits_l3.shape[1]) != len(l3_list):
        raise ValueError(f"logits_l3 has {int(logits_l3.shape[1])} classes, but l3_to_l2 has {len(l3_list)} entries.")
    if reduce not in {"sum", "mean", "logsumexp"}:
        raise ValueError(f"Unsupported reduce='{reduce}'. Expected one of: sum, mean, logsumexp.")
    if _fused_ok(logits_l3, num_l2) and all(0 <= int(v) < num_l2 for v in l3_list):
        from . import ops  # one launch instead of a Python loop of C3 slice updates
        return ops.l2_metrics(logits_l3, l3_list, num_l2, reduce, k=0, want_top3=False)[0]
    shape = (logits_l3.shape[0], num_l2)
    if reduce == "logsumexp":
        out = torch.full(shape, float("-inf"), device=logits_l3.device, dtype=logits_l3.dtype)
        for l3_id, l2_id in enumerate(l3_list):
            out[:, l2_id] = torch.logaddexp(out[:, l2_id], logits_l3[:, l3_id])
        return out
    out = torch.zeros(shape, device=logits_l3.device, dtype=logits_l3.dtype)
    counts = torch.zeros(num_l2, device=logits_l3.device, dtype=logits_l3.dtype)
    for l3_id, l2_id in enumerate(l3_list):
        out[:, l2_id] += logits_l3[:, l3_id]
        counts[l2_id] += 1
    return out / counts.clamp_min(1) if reduce == "mean" else out


def _weighted_f1(y_true: np.ndarray, y_pred: np.ndarray, num_classes: int) -> float:
    """Support-weighted mean of per-class F1 (torcheval MulticlassF1Score(average='weighted') semantics)."""
    f1s, weights = [], []
    for c in range(num_classes):
        tp = float(((y_pred == c) & (y_true == c)).sum())
        fp = float(((y_pred == c) & (y_true != c)).sum())
        fn = float(((y_pred != c) & (y_true == c)).sum())
        support = tp + fn
        if support == 0:
            continue
        f1s.append(2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else 0.0)
        weights.append(support)
    return float(np.average(f1s, weights=weights)) if f1s else 0.0


def _mcc(y_true: np.ndarray, y_pred: np.ndarray, num_classes: int) -> float:
    """Multiclass Matthews correlation coefficient (sklearn.metrics.matthews_corrcoef definition)."""
    cm = np.zeros((num_classes, num_classes), dtype=np.float64)
    np.add.at(cm, (y_true, y_pred), 1)
    t, p, c, s = cm.sum(1), cm.sum(0), np.trace(cm), cm.sum()
    cov_tp, cov_pp, cov_tt = c * s - t @ p, s * s - p @ p, s * s - t @ t
    return 0.0 if cov_pp * cov_tt == 0 else float(cov_tp / np.sqrt(cov_pp * cov_tt))


class L2MetricsAccumulator:
    """ref aihab_utils/evaluation.py:145-250 — top-k accuracy, weighted F1, MCC and optional confusion matrix on
    L2 labels from L3 logits.  mode 'argmax' maps the L3 argmax to L2 (top-1 only); mode 'logits' aggregates L3
    logits to L2 first."""

    def __init__(self, l3_to_l2, num_l2: int, reduce: str = "mean", topk: Sequence[int] = (1, 3),
                 return_confusion_matrix: bool = False, mode: str = "argmax") -> None:
        if mode not in {"argmax", "logits"}:
            raise ValueError(f"Unsupported mode='{mode}'. Expected 'argmax' or 'logits'.")
        self.l3_to_l2, self.num_l2, self.reduce, self.mode = l3_to_l2, int(num_l2), reduce, mode
        self.topk = (1,) if mode == "argmax" else tuple(int(k) for k in topk)
        self.return_confusion_matrix = return_confusion_matrix
        self.total_seen = 0
        self.correct_at_k = {k: 0 for k in self.topk}
        self.y_true_all, self.y_pred_all = [], []

    def update(self, logits_l3: torch.Tensor, targets_l3: torch.Tensor) -> None:
        targets_l2 = map_l3_targets_to_l2(targets_l3, self.l3_to_l2)
        batch = int(targets_l2.shape[0])
        self.total_seen += batch
        if batch == 0:
            return
        if self.mode == "argmax":
            preds = map_l3_targets_to_l2(logits_l3.argmax(dim=1), self.l3_to_l2)
            self.correct_at_k[1] += int((preds == targets_l2).sum().item())
        else:
            max_k = min(max(self.topk), self.num_l2)
            if _fused_ok(logits_l3, self.num_l2) and self.reduce in {"sum", "mean", "logsumexp"}:
                from . import ops  # aggregation + top-k in one launch
                logits_l2, top_idx, _, _, _ = ops.l2_metrics(logits_l3, self.l3_to_l2, self.num_l2, self.reduce,
                                                             k=max_k, want_top3=False)
            else:
                logits_l2 = aggregate_logits_to_l2(logits_l3, self.l3_to_l2, self.num_l2, reduce=self.reduce)
                top_idx = logits_l2.topk(max_k, dim=1).indices
            correct = top_idx.eq(targets_l2.view(-1, 1))
            for k in self.topk:
                k_eff = min(k, max_k)
                if k_eff >= 1:
                    self.correct_at_k[k] += int(correct[:, :k_eff].any(dim=1).sum().item())
            preds = logits_l2.argmax(dim=1)
        self.y_true_all.append(targets_l2.detach().cpu())
        self.y_pred_all.append(preds.detach().cpu())

    def compute(self) -> dict:
        denom = max(self.total_seen, 1)
        metrics = {f"top{k}": self.correct_at_k.get(k, 0) / denom for k in self.topk}
        if self.total_seen == 0:
            metrics.update(f1=0.0, mcc=0.0,
                           cm=np.zeros((self.num_l2, self.num_l2)) if self.return_confusion_matrix else None)
            return metrics
        y_true = torch.cat(self.y_true_all).numpy()
        y_pred = torch.cat(self.y_pred_all).numpy()
        metrics["f1"] = _weighted_f1(y_true, y_pred, self.num_l2)
        metrics["mcc"] = _mcc(y_true, y_pred, self.num_l2)
        if self.return_confusion_matrix:
            cm = np.zeros((self.num_l2, self.num_l2), dtype=np.int64)
            np.add.at(cm, (y_true, y_pred), 1)
            metrics["cm"] = cm
        else:
            metrics["cm"] = None
        return metrics


class ClassificationTracker:
    """ref aihab_utils/evaluation.py:253-273 (metrics part)."""

    def __init__(self) -> None:
        self.misclassified = []
        self.accurate_classified = []

    def top3_metrics(self, outputs: torch.Tensor, labels: torch.Tensor):
        if _fused_ok(outputs, min(int(outputs.shape[1]), 256)) and 3 <= outputs.shape[1] <= 256:
            from . import ops  # top-3 + softmax probabilities in one launch (identity L3 -> L2 map)
            _, _, _, top3_pred_indices, top3_probs = ops.l2_metrics(outputs, range(outputs.shape[1]), outputs.shape[1],
                                                                    "sum", k=0, want_logits=False)
        else:
            top3_pred_indices = torch.topk(outputs, 3, dim=1).indices
            top3_probs = torch.gather(F.softmax(outputs, dim=1), 1, top3_pred_indices)
        top3_correct = torch.sum(torch.any(top3_pred_indices == labels.unsqueeze(1), dim=1))
        return top3_correct, top3_pred_indices, top3_probs
