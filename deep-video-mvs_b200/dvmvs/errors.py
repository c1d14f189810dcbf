"""Depth error metrics with the reference's names, validity rule and order (dvmvs/errors.py:4-28), numpy.
Out of the hot path; present so that dvmvs.utils.save_results keeps its behaviour."""
import numpy as np


def compute_errors(gt, pred, max_depth=np.inf):
    valid = (gt >= 0.5) & (gt <= max_depth)            # errors.py:5-7
    g = np.asarray(gt)[valid].astype(np.float64)
    p = np.asarray(pred)[valid].astype(np.float64)
    if g.size == 0:
        return (np.nan,) * 8
    diff = np.abs(g - p)
    ratio = np.maximum(g / p, p / g)
    return (diff.mean(), (diff / g).mean(), np.abs(1.0 / g - 1.0 / p).mean(), (diff ** 2 / g).mean(),
            np.sqrt((diff ** 2).mean()), (ratio < 1.25).mean(), (ratio < 1.25 ** 2).mean(), (ratio < 1.25 ** 3).mean())
