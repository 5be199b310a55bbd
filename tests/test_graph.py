"""CPU tests: the vectorised graph construction vs the loop restatement of MIND_corpus.py:162-216."""
import numpy as np
import torch

from nnr_b200.synthetic import SyntheticMIND, history_structure
from oracle import graph as G


def test_history_structure_bit_exact():
    rng = np.random.default_rng(0)
    H, C = 50, 18
    cats = rng.integers(0, C, size=(200, H))
    hl = rng.integers(0, H + 1, size=200)
    hl[:3] = [0, 1, H]
    graph, mask, idx = history_structure(cats, hl, C)
    for b in range(200):
        g, m, i = G.build_history_graph(cats[b, :hl[b]], H, C)
        assert np.array_equal(g, graph[b].numpy()), b
        assert np.array_equal(m, mask[b].numpy()), b
        assert np.array_equal(i, idx[b].numpy()), b


def test_closed_form_degrees():
    g, m, i = G.build_history_graph([2, 2, 5, 2, 7], 8, 10)
    assert np.allclose(g.sum(1) > 0, True)
    # news of category 2: self + proxy + 2 peers = degree 4 -> d = 1/2
    assert g[0, 0] == np.float32(0.25)
    assert m.tolist() == [False, False, True, False, False, True, False, True, False, False, False]
    assert i.tolist() == [2, 2, 5, 2, 7, 10, 10, 10]


def test_synthetic_batch_contract():
    syn = SyntheticMIND(news_num=300, vocabulary_size=1000, max_history_num=50, lengths='mind', seed=1)
    b = syn.batch(4, seed=2)
    assert b['user_title_text'].shape == (4, 50, 32) and b['user_title_text'].dtype == torch.int32
    assert b['user_content_mask'].shape == (4, 50, 128) and b['user_content_mask'].dtype == torch.bool
    assert b['user_history_graph'].shape == (4, 68, 68) and b['user_history_graph'].dtype == torch.float32
    assert b['user_history_category_mask'].shape == (4, 19)
    assert b['user_history_category_indices'].dtype == torch.int64
    assert b['news_title_text'].shape == (4, 5, 32)
    # prefix masks, padded slots are the <PAD> news
    tm = b['user_title_mask']
    assert torch.all(tm[..., 1:] <= tm[..., :-1])
