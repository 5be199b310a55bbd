"""Device-resident news corpus and index-only batches.

The reference materialises every sample on the host -- 55 news x (32 + 128) token ids, masks, unused entity tensors
and an 18.5 KB dense history graph per impression (MIND_dataset.py:70-76, MIND_corpus.py:162-216) -- and copies
~99 KB per impression to the GPU each step.  Here the tokenised corpus (titles, abstracts, masks, category ids) is
uploaded ONCE; a training batch is ``history_ids [B,50]``, ``history_len [B]`` and ``candidate_ids [B,1+K]``
(or a positive id + the impression's negative pool, sampled on the device), the 21 model inputs are gathered on the
device, and the history graph / cluster mask / cluster indices are built by ``nnr_sue_graph_build`` (bit-exact with
MIND_corpus.py:178-213).  The gathered batch is bit-identical to the host-materialised one (tests/test_model_gpu.py).
"""
import torch

from . import ops


class DeviceCorpus:
    def __init__(self, news_title_text, news_title_mask, news_content_text, news_content_mask, news_category,
                 news_subCategory, category_num, device):
        to = lambda t, dt: torch.as_tensor(t).to(device=device, dtype=dt).contiguous()
        self.dev = torch.device(device)
        self.title_text, self.title_mask = to(news_title_text, torch.int32), to(news_title_mask, torch.bool)
        self.content_text, self.content_mask = to(news_content_text, torch.int32), to(news_content_mask, torch.bool)
        self.category, self.subCategory = to(news_category, torch.int32), to(news_subCategory, torch.int32)
        self.category_num = int(category_num)

    @classmethod
    def from_synthetic(cls, data, device):
        return cls(data.news_title_text, data.news_title_mask, data.news_abstract_text, data.news_abstract_mask,
                   data.news_category, data.news_subCategory, data.C, device)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.title_text, self.title_mask, self.content_text,
                                                          self.content_mask, self.category, self.subCategory))

    # ---- MIND_dataset.py:27-45 on the device ------------------------------------------------------------------
    def sample_negatives(self, negative_pool, pool_len, K, generator=None):
        """negative_pool [B, P] news ids (valid prefix of length pool_len[b] >= 1), -> [B, K] ids.
        pool_len <= K: cyclic ``pool[j % pool_len]`` exactly as MIND_dataset.py:34-36; otherwise K DISTINCT positions
        drawn uniformly (MIND_dataset.py:38-45 by rejection; here the K smallest of P random keys: same
        distribution, no loop)."""
        pool = torch.as_tensor(negative_pool).to(self.dev).long()
        plen = torch.as_tensor(pool_len).to(self.dev).long()
        B, P = pool.shape
        j = torch.arange(K, device=self.dev)[None, :]
        cyc = j % plen[:, None]
        keys = torch.rand(B, P, device=self.dev, generator=generator)
        keys = keys.masked_fill(torch.arange(P, device=self.dev)[None, :] >= plen[:, None], 2.0)
        rnd = torch.topk(keys, min(K, P), dim=1, largest=False)[1]
        if rnd.shape[1] < K:
            rnd = torch.cat([rnd, rnd.new_zeros(B, K - rnd.shape[1])], 1)
        pos = torch.where((plen <= K)[:, None], cyc, rnd)
        return torch.gather(pool, 1, pos)

    # ---- the 21 positional arguments of Model.forward (model.py:120-121), gathered on the device ----------------
    def batch_from_ids(self, history_ids, history_len, candidate_ids):
        hid = torch.as_tensor(history_ids).to(self.dev, non_blocking=True).long()
        hl = torch.as_tensor(history_len).to(self.dev, non_blocking=True).to(torch.int32)
        cid = torch.as_tensor(candidate_ids).to(self.dev, non_blocking=True).long()
        B, H = hid.shape
        C = self.category_num
        ucat = self.category[hid]
        graph = torch.empty(B, H + C, H + C, device=self.dev)
        cmask = torch.empty(B, C + 1, dtype=torch.bool, device=self.dev)
        cidx = torch.empty(B, H, dtype=torch.int64, device=self.dev)
        ops.sue_graph_build(ucat.contiguous(), hl, C, graph, cmask, cidx)
        hmask = torch.arange(H, device=self.dev)[None, :] < hl[:, None]
        return [torch.zeros(B, dtype=torch.int64, device=self.dev), ucat, self.subCategory[hid],
                self.title_text[hid], self.title_mask[hid], None, self.content_text[hid], self.content_mask[hid], None,
                hmask, graph, cmask, cidx,
                self.category[cid], self.subCategory[cid], self.title_text[cid], self.title_mask[cid], None,
                self.content_text[cid], self.content_mask[cid], None]
