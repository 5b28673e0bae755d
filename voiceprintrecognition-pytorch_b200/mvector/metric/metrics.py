"""Verification metrics used by MVectorTrainer.evaluate (reference: mvector/metric/metrics.py:5-43).  Pure numpy glue on
the score list; same definitions, argument meaning and return values as the reference:

  compute_fnr_fpr  thresholds = scores ascending; fnr[i] = share of target weight at or below thresholds[i],
                   fpr[i] = share of impostor weight strictly above it                    (metrics.py:5-20)
  compute_eer      linear interpolation between the last fnr < fpr point and the first fnr >= fpr point; with
                   ``scores`` also returns the score at the crossing index                  (metrics.py:23-33)
  compute_dcf      min over thresholds of the detection cost, normalised by the best trivial system (metrics.py:36-39)
"""
import numpy as np


def compute_fnr_fpr(scores, labels, weights=None):
    order = np.argsort(scores)
    thresholds = scores[order]
    lab = labels[order]
    wts = np.ones(lab.shape, dtype=np.float64) if weights is None else weights[order]
    tgt = np.where(lab == 1, wts, 0.0).astype(np.float64)
    imp = np.where(lab == 0, wts, 0.0).astype(np.float64)
    fnr = np.cumsum(tgt) / np.sum(tgt)
    fpr = 1 - np.cumsum(imp) / np.sum(imp)
    return fnr, fpr, thresholds


def compute_eer(fnr, fpr, scores=None):
    gap = fnr - fpr
    hi = np.flatnonzero(gap >= 0)[0]          # first threshold where misses overtake false accepts
    lo = np.flatnonzero(gap < 0)[-1]          # last threshold before that
    a = (fnr[hi] - fpr[hi]) / (fpr[lo] - fpr[hi] - (fnr[lo] - fnr[hi]))
    eer = fnr[hi] + a * (fnr[lo] - fnr[hi])
    if scores is None:
        return eer
    return eer, np.sort(scores)[hi]


def compute_dcf(fnr, fpr, p_target=0.01, c_miss=1, c_fa=1):
    cost = c_miss * p_target * fnr + c_fa * (1 - p_target) * fpr
    return cost.min() / min(c_miss * p_target, c_fa * (1 - p_target))
