"""Ranking + AUC / MRR / nDCG@5 / nDCG@10 on the device.

The reference writes per-impression ranks to a file on the host (util.py:52-62) and evaluates them with per-impression
numpy / sklearn calls (evaluate.py:31-88).  Here a whole padded batch of impressions [B, n_max] is ranked with one
stable descending sort and the four metrics come from prefix sums over the rank-ordered labels, in float64 like the
reference.  Ties keep the original candidate order, exactly as Python's stable ``list.sort(reverse=True)`` does.
"""
import torch


def rank_impressions(scores, counts):
    """scores [B, n_max] (entries beyond counts[b] ignored), counts [B] -> (ranks [B, n_max] int64, 1-based, 0 on the
    padding; order [B, n_max] = candidate index at each rank position)"""
    B, n = scores.shape
    valid = torch.arange(n, device=scores.device)[None, :] < counts[:, None]
    keyed = torch.where(valid, scores, torch.full_like(scores, float('-inf')))
    # padding must sort after every real candidate even if a real score is -inf: sort on (valid, score)
    order = torch.sort(keyed, dim=1, descending=True, stable=True)[1]
    order = torch.gather(order, 1, torch.sort(torch.gather(valid, 1, order).to(torch.int8), dim=1, descending=True, stable=True)[1])
    ranks = torch.zeros(B, n, dtype=torch.int64, device=scores.device)
    ranks.scatter_(1, order, torch.arange(1, n + 1, device=scores.device).expand(B, n).contiguous())
    return ranks * valid, order


def impression_metrics(scores, labels, counts):
    """Per-impression (auc, mrr, ndcg5, ndcg10), each [B] float64.  labels in {0, 1}; every impression needs at least
    one positive and one negative (sklearn's roc_auc_score raises otherwise in the reference)."""
    B, n = scores.shape
    dev = scores.device
    valid = torch.arange(n, device=dev)[None, :] < counts[:, None]
    _, order = rank_impressions(scores, counts)
    y = (torch.gather(labels.to(torch.float64), 1, order) * torch.gather(valid, 1, order)).contiguous()   # labels in rank order
    pos = torch.arange(n, device=dev, dtype=torch.float64)
    P = y.sum(1)
    N = counts.to(torch.float64) - P
    # AUC with distinct scores: fraction of (positive, negative) pairs with the positive ranked above
    neg_sorted = (1.0 - y) * torch.gather(valid, 1, order)
    neg_after = N[:, None] - torch.cumsum(neg_sorted, 1)
    auc = (y * neg_after).sum(1) / (P * N)
    mrr = (y / (pos + 1.0)).sum(1) / P
    gains = torch.pow(2.0, y) - 1.0
    disc = torch.log2(pos + 2.0)
    ideal = torch.sort(labels.to(torch.float64) * valid, dim=1, descending=True)[0]
    ig = torch.pow(2.0, ideal) - 1.0

    def ndcg(k):
        m = (pos < k).to(torch.float64)
        return (gains / disc * m).sum(1) / (ig / disc * m).sum(1)
    return auc, mrr, ndcg(5), ndcg(10)


def scoring(scores, labels, counts):
    """means over the impressions with counts > 0 (evaluate.py:45-46 skips empty ones) -> 4 python floats"""
    keep = counts > 0
    a, m, n5, n10 = impression_metrics(scores[keep], labels[keep], counts[keep])
    return float(a.mean()), float(m.mean()), float(n5.mean()), float(n10.mean())
