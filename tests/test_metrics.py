"""nnr_b200.metrics (device ranking + AUC/MRR/nDCG) against the restated evaluate.py / util.py (oracle/metrics.py)."""
import numpy as np
import pytest
import torch

from nnr_b200 import metrics
from oracle import metrics as OM


def _case(seed, B=37, nmax=60, ties=False):
    rng = np.random.default_rng(seed)
    counts = rng.integers(2, nmax + 1, size=B)
    scores = rng.normal(size=(B, nmax)).astype(np.float32)
    if ties:
        scores = np.round(scores * 2) / 2                      # many equal scores: stable-sort tie-breaking matters
    labels = np.zeros((B, nmax), dtype=np.int64)
    for b in range(B):
        k = rng.integers(1, max(2, counts[b] // 3))
        labels[b, rng.choice(counts[b], size=min(k, counts[b] - 1), replace=False)] = 1
    return scores, labels, counts


@pytest.mark.parametrize('ties', [False, True])
def test_ranks_and_metrics_match_reference_restatement(ties):
    scores, labels, counts = _case(3 + ties, ties=ties)
    ts, tl, tc = torch.from_numpy(scores), torch.from_numpy(labels), torch.from_numpy(counts)
    ranks, _ = metrics.rank_impressions(ts, tc)
    for b in range(len(counts)):
        assert ranks[b, :counts[b]].tolist() == OM.ranks_from_scores(scores[b, :counts[b]].tolist())   # bit-exact ranks
        assert ranks[b, counts[b]:].abs().sum() == 0
    ref = OM.scoring([labels[b, :counts[b]].tolist() for b in range(len(counts))],
                     [scores[b, :counts[b]].tolist() for b in range(len(counts))])
    got = metrics.scoring(ts, tl, tc)
    for r, g in zip(ref, got):
        assert abs(r - g) < 1e-12, (ref, got)


@pytest.mark.gpu
def test_metrics_on_device(cuda):
    scores, labels, counts = _case(11, B=64, nmax=40, ties=True)
    ref = OM.scoring([labels[b, :counts[b]].tolist() for b in range(len(counts))],
                     [scores[b, :counts[b]].tolist() for b in range(len(counts))])
    got = metrics.scoring(torch.from_numpy(scores).to(cuda), torch.from_numpy(labels).to(cuda), torch.from_numpy(counts).to(cuda))
    for r, g in zip(ref, got):
        assert abs(r - g) < 1e-12, (ref, got)
