"""Seeded synthetic MIND-shaped corpus, behaviours and batches (host side, numpy/torch CPU).

There is no network, so the MIND data the reference trains on (MIND_corpus.py:223-414) is replaced
by a generator with the same tensor contract (MIND_dataset.py:49-76): int32 token tables
``[news_num, 32]`` / ``[news_num, 128]`` with prefix bool masks, int32 category / subCategory,
behaviours = (history news ids padded with the <PAD> news 0, 1 + K candidate ids), and the
per-behaviour graph / category mask / cluster indices of MIND_corpus.py:162-216.

``history_structure`` is this package's own vectorised construction of that graph (closed form:
cliques per category + news<->proxy stars + clique over present proxies, then the symmetric
normalisation in fp32); it is checked bit for bit against the loop restatement in
``oracle/graph.py`` by ``tests/test_graph.py``.  The device version is ``nnr_sue_graph_build``.
"""
import numpy as np
import torch

FIELDS = ['user_ID', 'user_category', 'user_subCategory', 'user_title_text', 'user_title_mask',
          'user_title_entity', 'user_content_text', 'user_content_mask', 'user_content_entity',
          'user_history_mask', 'user_history_graph', 'user_history_category_mask',
          'user_history_category_indices', 'news_category', 'news_subCategory', 'news_title_text',
          'news_title_mask', 'news_title_entity', 'news_content_text', 'news_content_mask',
          'news_content_entity']


def history_structure(categories, history_len, category_num, normalize=True, no_self_connection=False,
                      gcn_normalization_type='symmetric'):
    """categories [B,H] int (category of each history slot, valid slots first), history_len [B].
    Returns graph [B,H+C,H+C] float32, category_mask [B,C+1] bool, category_indices [B,H] int64.
    Same values as MIND_corpus.py:178-213; ``normalize=False`` = no_adjacent_normalization,
    ``no_self_connection`` (:179-182) and ``gcn_normalization_type='asymmetric'`` (:204-208) are the reference's flags
    (defaults: self connections, symmetric normalisation).  As in the reference an empty history is never normalised."""
    cats = torch.as_tensor(categories).long()
    hl = torch.as_tensor(history_len).long()
    B, H = cats.shape
    C = int(category_num)
    G = H + C
    if no_self_connection and normalize:
        raise ValueError('adjacent normalisation needs self connections (reference config.py:111)')
    valid = torch.arange(H).unsqueeze(0) < hl.unsqueeze(1)                    # [B,H]
    idx = torch.where(valid, cats, torch.full_like(cats, C))                  # pad slots -> cluster C
    onehot = torch.zeros(B, H, C + 1, dtype=torch.float32).scatter_(2, idx.unsqueeze(2), 1.0)[:, :, :C]
    present = onehot.sum(1) > 0                                               # [B,C]
    mask = torch.zeros(B, C + 1, dtype=torch.bool)
    mask[:, :C] = present
    A = torch.zeros(B, G, G, dtype=torch.float32)
    A[:, :H, :H] = torch.bmm(onehot, onehot.transpose(1, 2))                  # same-category pairs (valid only)
    A[:, :H, H:] = onehot                                                     # news <-> proxy
    A[:, H:, :H] = onehot.transpose(1, 2)
    pf = present.float()
    A[:, H:, H:] = pf.unsqueeze(2) * pf.unsqueeze(1)                          # clique over present proxies
    eye = torch.eye(G, dtype=torch.float32).unsqueeze(0)
    if no_self_connection:
        A = A * (1.0 - eye)                                                   # the pair products above set the diagonal
    else:
        A = torch.maximum(A, eye.expand(B, -1, -1))                           # self connections
    if normalize:
        deg = A.sum(dim=2)                                                    # exact small integers
        nonempty = (hl > 0).view(B, 1, 1)
        if gcn_normalization_type == 'asymmetric':
            An = (1.0 / deg).unsqueeze(2) * A                                 # D^-1 A
        else:
            d = torch.sqrt(1.0 / deg)                                         # fp32, correctly rounded
            An = (d.unsqueeze(2) * A) * d.unsqueeze(1)
        A = torch.where(nonempty, An, A)
    return A, mask, idx


class SyntheticMIND:
    """Corpus tables + behaviour sampler.  All randomness comes from ``seed``."""

    def __init__(self, news_num=20000, vocabulary_size=40000, category_num=18, subCategory_num=285,
                 max_title_length=32, max_abstract_length=128, max_history_num=50,
                 negative_sample_num=4, lengths='mind', seed=0):
        self.news_num, self.V, self.C, self.S = news_num, vocabulary_size, category_num, subCategory_num
        self.T, self.A, self.H, self.K = max_title_length, max_abstract_length, max_history_num, negative_sample_num
        self.lengths = lengths
        rng = np.random.default_rng(seed)
        self.rng = rng
        n = news_num
        cat_p = rng.dirichlet(np.full(category_num, 0.6))
        self.news_category = rng.choice(category_num, size=n, p=cat_p).astype(np.int32)
        self.news_subCategory = rng.integers(1, subCategory_num, size=n).astype(np.int32)
        if lengths == 'full':
            tl = np.full(n, self.T)
            al = np.full(n, self.A)
        elif lengths == 'uniform':
            tl = rng.integers(4, self.T + 1, size=n)
            al = rng.integers(8, self.A + 1, size=n)
        else:  # 'mind': title ~ N(12,4), abstract ~ lognormal with mean ~40
            tl = np.clip(np.rint(rng.normal(12.0, 4.0, size=n)), 1, self.T).astype(np.int64)
            al = np.clip(np.rint(rng.lognormal(np.log(40.0) - 0.32, 0.8, size=n)), 1, self.A).astype(np.int64)
        self.title_len, self.abstract_len = tl.astype(np.int64), al.astype(np.int64)
        # Zipf(~1) over [2, V): inverse-CDF sampling on a truncated harmonic law
        ranks = np.arange(1, vocabulary_size - 1, dtype=np.float64)
        cdf = np.cumsum(1.0 / ranks)
        cdf /= cdf[-1]

        def zipf_ids(shape):
            u = rng.random(size=shape)
            return (np.searchsorted(cdf, u) + 2).astype(np.int32)

        self.news_title_text = zipf_ids((n, self.T))
        self.news_abstract_text = zipf_ids((n, self.A))
        self.news_title_mask = np.arange(self.T)[None, :] < tl[:, None]
        self.news_abstract_mask = np.arange(self.A)[None, :] < al[:, None]
        self.news_title_text[~self.news_title_mask] = 0
        self.news_abstract_text[~self.news_abstract_mask] = 0
        # <PAD> news 0: no words, mask only [0] set (MIND_corpus.py:352-353), category 0
        self.news_title_text[0] = 0
        self.news_abstract_text[0] = 0
        self.news_title_mask[0] = False
        self.news_abstract_mask[0] = False
        self.news_title_mask[0, 0] = True
        self.news_abstract_mask[0, 0] = True
        self.news_category[0] = 0
        self.news_subCategory[0] = 0
        self.title_len[0] = 1
        self.abstract_len[0] = 1
        # news grouped by category for user-skewed sampling
        self._by_cat = [np.nonzero((self.news_category == c) & (np.arange(n) > 0))[0] for c in range(category_num)]

    def word_table(self, dim=300, seed=1):
        """N(0, 0.3^2) rows, row 0 (<PAD>) = 0 -- stands in for GloVe (MIND_corpus.py:114-132)."""
        g = torch.Generator().manual_seed(seed)
        w = torch.randn(self.V, dim, generator=g) * 0.3
        w[0].zero_()
        return w

    def sample_behaviors(self, batch_size, news_num=None, seed=0):
        """Returns (history ids [B,H] with 0 padding at the end, history_len [B], candidates [B,n])."""
        rng = np.random.default_rng(seed + 7919)
        n = (1 + self.K) if news_num is None else news_num
        hist = np.zeros((batch_size, self.H), dtype=np.int64)
        hl = np.where(rng.random(batch_size) < 0.1, self.H, rng.integers(0, self.H + 1, size=batch_size))
        for b in range(batch_size):
            p = rng.dirichlet(np.full(self.C, 0.3))
            cats = rng.choice(self.C, size=int(hl[b]), p=p)
            for i, c in enumerate(cats):
                pool = self._by_cat[c]
                hist[b, i] = pool[rng.integers(0, len(pool))] if len(pool) else rng.integers(1, self.news_num)
        cand = rng.integers(1, self.news_num, size=(batch_size, n)).astype(np.int64)
        return hist, hl.astype(np.int64), cand

    def batch(self, batch_size, news_num=None, seed=0):
        """The reference's 21-tensor batch (MIND_dataset.py:70-76 dtypes/shapes) as a dict of CPU
        tensors, plus 'history_ids', 'history_len', 'candidate_ids' for the index-only path."""
        hist, hl, cand = self.sample_behaviors(batch_size, news_num, seed)
        return self.materialize(hist, hl, cand)

    def materialize(self, hist, hl, cand):
        t = torch.from_numpy
        B = hist.shape[0]
        ucat = self.news_category[hist]
        graph, cmask, cidx = history_structure(ucat, hl, self.C)
        out = {
            'user_ID': torch.zeros(B, dtype=torch.int64),
            'user_category': t(ucat.astype(np.int32)),
            'user_subCategory': t(self.news_subCategory[hist].astype(np.int32)),
            'user_title_text': t(self.news_title_text[hist]),
            'user_title_mask': t(self.news_title_mask[hist]),
            'user_title_entity': None,
            'user_content_text': t(self.news_abstract_text[hist]),
            'user_content_mask': t(self.news_abstract_mask[hist]),
            'user_content_entity': None,
            'user_history_mask': t(np.arange(self.H)[None, :] < hl[:, None]),
            'user_history_graph': graph,
            'user_history_category_mask': cmask,
            'user_history_category_indices': cidx,
            'news_category': t(self.news_category[cand].astype(np.int32)),
            'news_subCategory': t(self.news_subCategory[cand].astype(np.int32)),
            'news_title_text': t(self.news_title_text[cand]),
            'news_title_mask': t(self.news_title_mask[cand]),
            'news_title_entity': None,
            'news_content_text': t(self.news_abstract_text[cand]),
            'news_content_mask': t(self.news_abstract_mask[cand]),
            'news_content_entity': None,
            'history_ids': t(hist), 'history_len': t(hl), 'candidate_ids': t(cand),
        }
        return out


def batch_args(batch, device=None):
    """The 21 positional arguments of Model.forward in reference order (model.py:120-121)."""
    args = []
    for k in FIELDS:
        v = batch[k]
        if torch.is_tensor(v) and device is not None:
            v = v.to(device, non_blocking=True)
        args.append(v)
    return args
