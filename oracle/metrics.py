"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dev/test ranking and metrics.

Follows util.py:52-62 (per-impression ranks: sort the scores descending with Python's stable sort, rank = position+1)
and evaluate.py:7-27,66-88 (y_score = 1/rank; AUC = sklearn.metrics.roc_auc_score; MRR; nDCG@5; nDCG@10; means over
impressions).  numpy float64 exactly as the reference computes them.  Imported by tests/ only.
"""
import numpy as np


def ranks_from_scores(scores):
    """util.py:55-61: stable descending sort; result[original position] = rank (1-based)"""
    sub = [[float(s), i] for i, s in enumerate(scores)]
    sub.sort(key=lambda x: x[0], reverse=True)
    result = [0] * len(sub)
    for j, (_, i) in enumerate(sub):
        result[i] = j + 1
    return result


def dcg_score(y_true, y_score, k=10):            # evaluate.py:7-12
    order = np.argsort(y_score)[::-1]
    y_true = np.take(y_true, order[:k])
    gains = 2 ** y_true - 1
    discounts = np.log2(np.arange(len(y_true)) + 2)
    return np.sum(gains / discounts)


def ndcg_score(y_true, y_score, k=10):           # evaluate.py:15-18
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true, k)


def mrr_score(y_true, y_score):                  # evaluate.py:21-25
    order = np.argsort(y_score)[::-1]
    y_true = np.take(y_true, order)
    rr_score = y_true / (np.arange(len(y_true)) + 1)
    return np.sum(rr_score) / np.sum(y_true)


def auc_score(y_true, y_score):
    """sklearn.metrics.roc_auc_score (evaluate.py:5,76) when importable; the scores are 1/rank, i.e. distinct, so the
    Mann-Whitney count below is the same number"""
    try:
        from sklearn.metrics import roc_auc_score
        return float(roc_auc_score(y_true, y_score))
    except ImportError:                            # pragma: no cover
        pos = [s for s, y in zip(y_score, y_true) if y > 0]
        neg = [s for s, y in zip(y_score, y_true) if y <= 0]
        return sum((p > n) + 0.5 * (p == n) for p in pos for n in neg) / (len(pos) * len(neg))


def scoring(labels_per_impression, scores_per_impression):
    """evaluate.py:31-88 without the file parsing: lists of per-impression label / score lists -> 4 means"""
    aucs, mrrs, n5, n10 = [], [], [], []
    for labels, scores in zip(labels_per_impression, scores_per_impression):
        if len(labels) == 0:
            continue
        ranks = ranks_from_scores(scores)
        y_true = np.array(labels, dtype='float32')
        y_score = [1. / r for r in ranks]
        aucs.append(auc_score(y_true, y_score))
        mrrs.append(mrr_score(y_true, y_score))
        n5.append(ndcg_score(y_true, y_score, 5))
        n10.append(ndcg_score(y_true, y_score, 10))
    return np.mean(aucs), np.mean(mrrs), np.mean(n5), np.mean(n10)
