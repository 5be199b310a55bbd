"""Cached-corpus dev/test scoring (BASELINE config 3).

The reference's ``util.compute_scores`` re-encodes the user's 50 history news for every (impression,
candidate) pair (util.py:19-50, README.md:125).  Here the news corpus is encoded ONCE with the CNE kernels
and kept resident in HBM as a [news_num, 900] table; an impression is then just 50 history ids + n candidate
ids: the history graph / category mask / cluster indices are built on the device from the cached category ids
(``nnr_sue_graph_build``, bit-exact with MIND_corpus.py:178-213), followed by SUE and the dot product.

Parity note (SURVEY finding 2): a news vector depends on the composition of the ``news_encoder`` call it is
encoded in (sort-rank pairing of the selective gate).  The cache is therefore defined per corpus chunk: news
[c*chunk, (c+1)*chunk) are encoded as ONE call of shape [1, chunk]; the oracle test encodes the same chunks.
"""
import torch

from . import engine, metrics, ops


class CorpusScorer:
    def __init__(self, model, news_title_text, news_title_mask, news_content_text, news_content_mask, news_category,
                 news_subCategory, chunk=4096):
        self.model = model
        self.dev = next(model.parameters()).device
        to = lambda t: torch.as_tensor(t).to(self.dev)
        self.tt, self.tm, self.ct, self.cm = to(news_title_text), to(news_title_mask).clone(), to(news_content_text), to(news_content_mask).clone()
        self.cat, self.sub = to(news_category).to(torch.int32), to(news_subCategory).to(torch.int32)
        self.chunk = chunk
        self.cache = None
        self._graphs = {}

    def _graphed(self, key, fn, example_inputs):
        """capture ``fn(*static_inputs)`` once per shape key and replay it: an encode of one corpus chunk is ~100 kernels of a
        few microseconds to a millisecond, i.e. host-launch bound when issued one by one (eval mode: no dropout, no seeds)"""
        g = self._graphs.get(key)
        if g is None:
            static = [x.clone() for x in example_inputs]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                           # warm-up outside capture: lazy attributes, workspaces, weight planes
                    fn(*[x.clone() for x in static])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn(*static)
            g = (static, graph, out)
            self._graphs[key] = g
        static, graph, out = g
        for s_, x in zip(static, example_inputs):
            s_.copy_(x, non_blocking=True)
        graph.replay()
        return out

    @torch.no_grad()
    def encode_corpus(self, cuda_graph=True):
        """news table -> [news_num, news_embedding_dim] fp32 cache; returns it.  Full chunks replay one captured graph."""
        self.model.eval()
        enc = self.model.news_encoder
        n = self.tt.shape[0]
        out = torch.empty(n, enc.news_embedding_dim, device=self.dev)

        def encode(tt, tm, ct, cm, cat, sub):
            return enc(tt.unsqueeze(0), tm.unsqueeze(0), None, ct.unsqueeze(0), cm.unsqueeze(0), None, cat.unsqueeze(0), sub.unsqueeze(0), None)[0]
        for a in range(0, n, self.chunk):
            b = min(n, a + self.chunk)
            args = (self.tt[a:b], self.tm[a:b], self.ct[a:b], self.cm[a:b], self.cat[a:b], self.sub[a:b])
            if cuda_graph and b - a == self.chunk and n >= 3 * self.chunk:
                out[a:b] = self._graphed(('encode', self.chunk), encode, args)
            else:
                out[a:b] = encode(*args)
        self.cache = out
        return out

    @torch.no_grad()
    def score(self, history_ids, history_len, candidate_ids, cuda_graph=True):
        """history_ids [B,H] (0-padded at the end), history_len [B], candidate_ids [B,n] -> scores [B,n]"""
        assert self.cache is not None, 'call encode_corpus() first'
        hid = torch.as_tensor(history_ids).to(self.dev).long()
        cid = torch.as_tensor(candidate_ids).to(self.dev).long()
        hl = torch.as_tensor(history_len).to(self.dev).to(torch.int32)
        if cuda_graph and hid.shape[0] >= 64:
            key = ('score', tuple(hid.shape), tuple(cid.shape), self.cache.data_ptr())
            return self._graphed(key, self._score_impl, (hid, hl, cid)).clone()
        return self._score_impl(hid, hl, cid)

    def _score_impl(self, hid, hl, cid):
        ue = self.model.user_encoder
        B, H = hid.shape
        C = ue.proxy_node_embedding.shape[0]
        cats = self.cat[hid].contiguous()
        graph = torch.empty(B, H + C, H + C, device=self.dev)
        cmask = torch.empty(B, C + 1, dtype=torch.bool, device=self.dev)
        cidx = torch.empty(B, H, dtype=torch.int64, device=self.dev)
        ops.sue_graph_build(cats, hl, C, graph, cmask, cidx)
        hist = self.cache[hid]                                  # [B, H, D]
        cand = self.cache[cid]                                  # [B, n, D]
        user = ue.encode_user(hist, graph, cmask, cidx, cand)
        return engine.RowDot.apply(user, cand)

    @torch.no_grad()
    def evaluate(self, history_ids, history_len, candidate_ids, candidate_count, labels, batch=256):
        """Dev-set metrics with the cached corpus and on-device ranking (replaces util.compute_scores +
        evaluate.scoring): impressions are padded to [I, n_max] candidate ids with ``candidate_count`` valid entries
        and 0/1 ``labels``.  Returns (auc, mrr, ndcg5, ndcg10, ranks [I, n_max])."""
        cid = torch.as_tensor(candidate_ids)
        out = []
        for a in range(0, cid.shape[0], batch):
            out.append(self.score(history_ids[a:a + batch], history_len[a:a + batch], cid[a:a + batch]))
        scores = torch.cat(out)
        cnt = torch.as_tensor(candidate_count).to(self.dev)
        lab = torch.as_tensor(labels).to(self.dev)
        ranks, _ = metrics.rank_impressions(scores, cnt)
        return metrics.scoring(scores, lab, cnt) + (ranks,)
