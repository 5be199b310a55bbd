"""TEST INFRASTRUCTURE -- plain-PyTorch CPU restatement of the reference CNE+SUE hot path.

This module is the *checker* for the CUDA path (and, in ``bench.py``, the timed CPU baseline).
It is never imported by ``nnr_b200``.  Every function cites the reference lines it restates
(paths relative to the reference checkout).  It is written functionally over a flat dict of
parameters that uses the reference's ``state_dict`` key names, so the same weights can be loaded
into the reference, this oracle and the CUDA modules.

Pinned against the reference itself (see ``oracle/__init__.py``): ``tests/golden`` holds outputs of
the real ``model.Model`` and ``tests/test_oracle.py`` replays them here.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# config / parameters
# ----------------------------------------------------------------------------------------------
def make_config(**kw):
    """Stub of the attributes CNE/SUE/Model read from reference ``config.py:28-76`` (defaults)."""
    cfg = dict(word_embedding_dim=300, vocabulary_size=1000, word_threshold=3, tokenizer='MIND',
               max_title_length=32, max_abstract_length=128, dataset='synthetic', category_num=18,
               subCategory_num=285, category_embedding_dim=50, subCategory_embedding_dim=50,
               dropout_rate=0.2, hidden_dim=200, attention_dim=200, gcn_layer_num=4,
               no_gcn_residual=False, gcn_layer_norm=False, max_history_num=50, news_encoder='CNE',
               user_encoder='SUE', click_predictor='dot_product', negative_sample_num=4,
               user_num=1, user_embedding_dim=50, batch_size=64, lr=1e-4, weight_decay=0,
               gradient_clip_norm=4.0, world_size=1)
    cfg.update(kw)
    return SimpleNamespace(**cfg)


def cne_modalities(cfg):
    return {'CNE_Title': ('title',), 'CNE_Content': ('content',)}.get(cfg.news_encoder, ('title', 'content'))


def param_shapes(cfg):
    """Unique parameter tensors of Model(CNE, SUE) with their reference checkpoint names
    (newsEncoders.py:12-77, userEncoders.py:43-56, layers.py:151-155,178-183,265-311).
    Variants: CNE_wo_CA has no cross-attention (variantEncoders.py:263-281); SUE_wo_HCA replaces
    the cluster attention by ``attention`` (variantEncoders.py:393-400)."""
    E, Hd, A = cfg.word_embedding_dim, cfg.hidden_dim, cfg.attention_dim
    mods = cne_modalities(cfg)                          # CNE_Title / CNE_Content: one modality (variantEncoders.py:14-99)
    D = 2 * Hd * len(mods) + cfg.category_embedding_dim + cfg.subCategory_embedding_dim
    Au = max(A, D // 4)
    s = {}
    ne = 'news_encoder.'
    s[ne + 'word_embedding.weight'] = (cfg.vocabulary_size, E)
    s[ne + 'category_embedding.weight'] = (cfg.category_num, cfg.category_embedding_dim)
    s[ne + 'subCategory_embedding.weight'] = (cfg.subCategory_num, cfg.subCategory_embedding_dim)
    for x in mods:
        for sfx in ('', '_reverse'):
            s[ne + f'{x}_lstm.weight_ih_l0{sfx}'] = (4 * Hd, E)
            s[ne + f'{x}_lstm.weight_hh_l0{sfx}'] = (4 * Hd, Hd)
            s[ne + f'{x}_lstm.bias_ih_l0{sfx}'] = (4 * Hd,)
            s[ne + f'{x}_lstm.bias_hh_l0{sfx}'] = (4 * Hd,)
    if cfg.news_encoder in ('CNE', 'CNE_wo_CA'):        # the other variants have no gate parameters
        for x in ('title', 'content'):
            s[ne + f'{x}_H.weight'] = (2 * Hd, 2 * Hd)
            s[ne + f'{x}_M.weight'] = (2 * Hd, 2 * Hd)
            s[ne + f'{x}_M.bias'] = (2 * Hd,)
    for x in mods:
        s[ne + f'{x}_self_attention.affine1.weight'] = (A, 2 * Hd)
        s[ne + f'{x}_self_attention.affine1.bias'] = (A,)
        s[ne + f'{x}_self_attention.affine2.weight'] = (1, A)
    if cfg.news_encoder in ('CNE', 'CNE_wo_CS'):
        for x in ('title', 'content'):
            s[ne + f'{x}_cross_attention.K.weight'] = (A, 2 * Hd)
            s[ne + f'{x}_cross_attention.Q.weight'] = (A, 2 * Hd)
            s[ne + f'{x}_cross_attention.Q.bias'] = (A,)
    ue = 'user_encoder.'
    if cfg.user_encoder != 'SUE_wo_GCN':                # variantEncoders.py:342-353: no proxy nodes, no GCN
        s[ue + 'proxy_node_embedding'] = (cfg.category_num, D)
        for l in range(cfg.gcn_layer_num):
            s[ue + f'gcn.gcn_layers.{l}.W.weight'] = (D, D)
            s[ue + f'gcn.gcn_layers.{l}.W.bias'] = (D,)
            if getattr(cfg, 'gcn_layer_norm', False):   # layers.py:274-275
                s[ue + f'gcn.gcn_layers.{l}.layer_normalization.weight'] = (D,)
                s[ue + f'gcn.gcn_layers.{l}.layer_normalization.bias'] = (D,)
    if cfg.user_encoder in ('SUE', 'SUE_wo_GCN'):
        s[ue + 'intraCluster_K.weight'] = (Au, D)
        if cfg.user_encoder == 'SUE_wo_GCN':
            s[ue + 'intraCluster_K.bias'] = (Au,)
        s[ue + 'intraCluster_Q.weight'] = (Au, D)
        s[ue + 'intraCluster_Q.bias'] = (Au,)
        s[ue + 'clusterFeatureAffine.weight'] = (D, D)
        s[ue + 'clusterFeatureAffine.bias'] = (D,)
        s[ue + 'interClusterAttention.K.weight'] = (Au, D)
        s[ue + 'interClusterAttention.Q.weight'] = (Au, D)
        s[ue + 'interClusterAttention.Q.bias'] = (Au,)
    else:  # SUE_wo_HCA
        s[ue + 'attention.affine1.weight'] = (A, D)
        s[ue + 'attention.affine1.bias'] = (A,)
        s[ue + 'attention.affine2.weight'] = (1, A)
    return s


def formula_params(cfg, dtype=torch.float32, salt=0):
    """Deterministic, RNG-free parameter values (a closed formula of the element index) so that
    golden fixtures need not store ~20 M weights and do not depend on torch's RNG stream.
    Magnitudes mimic ``initialize()`` (orthogonal / xavier scale ~ 1/sqrt(fan)); biases are made
    non-zero on purpose so that every bias path is exercised."""
    out = {}
    for t, (name, shape) in enumerate(param_shapes(cfg).items()):
        n = 1
        for d in shape:
            n *= d
        i = torch.arange(n, dtype=torch.float64)
        v = torch.sin(i * (0.37 + 0.011 * ((t + salt) % 17)) + 1.3 * (t + salt)) \
            * torch.cos(i * 0.0123 + 0.7 * t)
        if len(shape) == 2:
            scale = 1.6 / math.sqrt(shape[1]) if 'embedding' not in name else 0.3
            if name.endswith('proxy_node_embedding'):
                scale = 0.05
            if 'category_embedding' in name or 'subCategory_embedding' in name:
                scale = 0.1
        else:
            scale = 0.05
        if name.endswith('layer_normalization.weight'):       # LayerNorm gain around 1
            v, scale = 1.0 + 0.1 * v, 1.0
        out[name] = (v * scale).reshape(shape).to(dtype)
    out['news_encoder.word_embedding.weight'][0].zero_()        # <PAD> row (MIND_corpus.py:121-124)
    return out


def alias_state_dict(params):
    """Reference state_dict also holds every news_encoder.* tensor under
    user_encoder.news_encoder.* (userEncoders.py:16 keeps the encoder as a sub-module)."""
    sd = dict(params)
    for k, v in params.items():
        if k.startswith('news_encoder.'):
            sd['user_encoder.' + k] = v
    return sd


# ----------------------------------------------------------------------------------------------
# torch_scatter 2.0.9 restatement (third-party, absent; call sites userEncoders.py:88-89)
# ----------------------------------------------------------------------------------------------
def _expand_index(index, src, dim):
    if index.dim() != src.dim():
        for _ in range(src.dim() - index.dim()):
            index = index.unsqueeze(-1)
    return index.expand_as(src)


def scatter_sum(src, index, dim, dim_size):
    """out[..., g, ...] = sum of src entries whose index == g (zeros elsewhere)."""
    index = _expand_index(index, src, dim)
    shape = list(src.shape)
    shape[dim] = dim_size
    return torch.zeros(shape, dtype=src.dtype).scatter_add_(dim, index, src)


def scatter_softmax(src, index, dim):
    """Per-group softmax: subtract the group max, exponentiate, divide by the group sum."""
    index = _expand_index(index, src, dim)
    size = int(index.max()) + 1
    shape = list(src.shape)
    shape[dim] = size
    gmax = torch.full(shape, float('-inf'), dtype=src.dtype).scatter_reduce(dim, index, src, 'amax', include_self=True)
    e = (src - gmax.gather(dim, index)).exp()
    gsum = torch.zeros(shape, dtype=src.dtype).scatter_add_(dim, index, e)
    return e / gsum.gather(dim, index)


# ----------------------------------------------------------------------------------------------
# CNE (newsEncoders.py:57-141)
# ----------------------------------------------------------------------------------------------
def lengths_and_perms(title_mask, content_mask, sort_fn=None):
    """newsEncoders.py:106-115.  Masks are [N, L] bool and are modified in place like the
    reference (column 0 forced to 1).  Returns lengths and the four permutations."""
    sort_fn = sort_fn or torch.sort
    title_mask[:, 0] = 1
    content_mask[:, 0] = 1
    tl = title_mask.sum(dim=1).long()
    cl = content_mask.sum(dim=1).long()
    _, st = sort_fn(tl, descending=True)
    _, dt = sort_fn(st, descending=False)
    _, sc = sort_fn(cl, descending=True)
    _, dc = sort_fn(sc, descending=False)
    return tl, cl, st, dt, sc, dc


_TORCH_SORT = torch.sort      # captured at import: tests patch torch.sort with stable_sort


def stable_sort(x, descending=False):
    return _TORCH_SORT(x, descending=descending, stable=True)


def lstm_direction(x, lens, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of a 1-layer nn.LSTM over variable-length rows (packed semantics,
    newsEncoders.py:119-127): each row runs exactly ``len`` steps from zero state; outputs beyond
    the length are zero; c_n is the cell state after the row's last step.  Gate order i,f,g,o."""
    N, L, _ = x.shape
    Hd = w_hh.shape[1]
    h = x.new_zeros(N, Hd)
    c = x.new_zeros(N, Hd)
    gx = x @ w_ih.t() + (b_ih + b_hh)
    outs = [None] * L
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        z = gx[:, t] + h @ w_hh.t()
        i, f, g, o = z.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        act = (t < lens).unsqueeze(1)
        c = torch.where(act, c_new, c)
        h = torch.where(act, h_new, h)
        outs[t] = torch.where(act, h_new, torch.zeros_like(h_new))
    return torch.stack(outs, dim=1), c


def bilstm(p, prefix, x, lens, impl='loop'):
    """Bidirectional LSTM -> (h [N,L,2Hd] with zeros beyond len, m = [c_n fwd | c_n bwd] [N,2Hd]).
    impl='aten' runs torch's own packed nn.LSTM path exactly like the reference (used for the
    timed CPU baseline and to cross-check the explicit loop)."""
    if impl == 'aten':
        from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
        E, Hd = x.shape[2], p[prefix + 'weight_hh_l0'].shape[1]
        m = torch.nn.LSTM(E, Hd, batch_first=True, bidirectional=True).to(x.dtype)
        with torch.no_grad():
            for k, v in m.named_parameters():
                v.copy_(p[prefix + k])
        # keep autograd connection to p: functional call
        pk = pack_padded_sequence(x, lens.cpu(), batch_first=True, enforce_sorted=False)
        names = [k for k, _ in m.named_parameters()]
        out, (h_n, c_n) = torch.func.functional_call(m, {k: p[prefix + k] for k in names}, (pk,))
        h, _ = pad_packed_sequence(out, batch_first=True, total_length=x.shape[1])
        return h, torch.cat([c_n[0], c_n[1]], dim=1)
    hf, cf = lstm_direction(x, lens, p[prefix + 'weight_ih_l0'], p[prefix + 'weight_hh_l0'],
                            p[prefix + 'bias_ih_l0'], p[prefix + 'bias_hh_l0'], False)
    hb, cb = lstm_direction(x, lens, p[prefix + 'weight_ih_l0_reverse'], p[prefix + 'weight_hh_l0_reverse'],
                            p[prefix + 'bias_ih_l0_reverse'], p[prefix + 'bias_hh_l0_reverse'], True)
    return torch.cat([hf, hb], dim=2), torch.cat([cf, cb], dim=1)


def additive_attention(p, prefix, feature, mask=None):
    """layers.py:167-175."""
    u = torch.tanh(feature @ p[prefix + 'affine1.weight'].t() + p[prefix + 'affine1.bias'])
    a = (u @ p[prefix + 'affine2.weight'].t()).squeeze(2)
    if mask is not None:
        a = a.masked_fill(mask == 0, -1e9)
    alpha = F.softmax(a, dim=1)
    return torch.bmm(alpha.unsqueeze(1), feature).squeeze(1)


def scaled_dot_candidate_attention(p, prefix, feature, query, mask=None):
    """layers.py:196-203."""
    k = feature @ p[prefix + 'K.weight'].t()
    q = query @ p[prefix + 'Q.weight'].t() + p[prefix + 'Q.bias']
    a = torch.bmm(k, q.unsqueeze(2)).squeeze(2) / math.sqrt(float(k.shape[2]))
    if mask is not None:
        a = a.masked_fill(mask == 0, -1e9)
    alpha = F.softmax(a, dim=1)
    return torch.bmm(alpha.unsqueeze(1), feature).squeeze(1)


def cne_forward(p, cfg, title_text, title_mask, content_text, content_mask, category, subCategory,
                pre='news_encoder.', sort_fn=None, lstm_impl='loop', dropout_masks=None,
                cross_attention=True, gate=True, modalities=('title', 'content')):
    """newsEncoders.py:102-141 (eval mode, or train mode with externally supplied keep-masks).

    The reference runs the LSTM in length-sorted order and adds the *other* modality's memory
    vector row by row in that order (:124-129), which pairs title row r (news st[r]) with the cell
    state of news sc[r].  Here everything stays in natural order and the pairing is an explicit
    gather: partner_title[i] = sc[dt[i]], partner_content[i] = st[dc[i]].
    """
    B, n = title_text.shape[0], title_text.shape[1]
    N = B * n
    T, A_len = cfg.max_title_length, cfg.max_abstract_length
    tm = title_mask.view(N, T)
    cm = content_mask.view(N, A_len)
    tl, cl, st, dt, sc, dc = lengths_and_perms(tm, cm, sort_fn)
    table = p[pre + 'word_embedding.weight']
    if len(modalities) == 1:                          # CNE_Title / CNE_Content (variantEncoders.py:35-55, 79-99)
        x = modalities[0]
        text, L, mask, ln = (title_text, T, tm, tl) if x == 'title' else (content_text, A_len, cm, cl)
        emb = table[text.reshape(N, L).long()]
        if dropout_masks is not None:
            emb = emb * dropout_masks[x]
        h, _ = bilstm(p, pre + x + '_lstm.', emb, ln, lstm_impl)
        rep = additive_attention(p, pre + x + '_self_attention.', h, mask).view(B, n, -1)
        cat_e = p[pre + 'category_embedding.weight'][category.long()]
        sub_e = p[pre + 'subCategory_embedding.weight'][subCategory.long()]
        if dropout_masks is not None:
            cat_e = cat_e * dropout_masks['category']
            sub_e = sub_e * dropout_masks['subCategory']
        return torch.cat([rep, cat_e, sub_e], dim=2)
    title = table[title_text.reshape(N, T).long()]
    content = table[content_text.reshape(N, A_len).long()]
    if dropout_masks is not None:                     # nn.Dropout: x * keep / (1 - p)
        title = title * dropout_masks['title']
        content = content * dropout_masks['content']
    th, t_m = bilstm(p, pre + 'title_lstm.', title, tl, lstm_impl)
    ch, c_m = bilstm(p, pre + 'content_lstm.', content, cl, lstm_impl)
    if gate:                                          # CNE_wo_CS (variantEncoders.py:244-252) skips the selective gate
        c_m_for_title = c_m[sc[dt]]                   # content memory paired by sort rank
        t_m_for_content = t_m[st[dc]]
        t_gate = torch.sigmoid(th @ p[pre + 'title_H.weight'].t()
                               + (c_m_for_title @ p[pre + 'title_M.weight'].t() + p[pre + 'title_M.bias']).unsqueeze(1))
        c_gate = torch.sigmoid(ch @ p[pre + 'content_H.weight'].t()
                               + (t_m_for_content @ p[pre + 'content_M.weight'].t() + p[pre + 'content_M.bias']).unsqueeze(1))
        th = th * t_gate
        ch = ch * c_gate
    t_self = additive_attention(p, pre + 'title_self_attention.', th, tm)
    c_self = additive_attention(p, pre + 'content_self_attention.', ch, cm)
    if cross_attention:
        t_cross = scaled_dot_candidate_attention(p, pre + 'title_cross_attention.', th, c_self, tm)
        c_cross = scaled_dot_candidate_attention(p, pre + 'content_cross_attention.', ch, t_self, cm)
        rep = torch.cat([t_self + t_cross, c_self + c_cross], dim=1)
    else:                                             # CNE_wo_CA, variantEncoders.py:336
        rep = torch.cat([t_self, c_self], dim=1)
    rep = rep.view(B, n, -1)
    cat_e = p[pre + 'category_embedding.weight'][category.long()]          # newsEncoders.py:50-54
    sub_e = p[pre + 'subCategory_embedding.weight'][subCategory.long()]
    if dropout_masks is not None:
        cat_e = cat_e * dropout_masks['category']
        sub_e = sub_e * dropout_masks['subCategory']
    return torch.cat([rep, cat_e, sub_e], dim=2)


# ----------------------------------------------------------------------------------------------
# SUE (userEncoders.py:42-98, layers.py:265-323)
# ----------------------------------------------------------------------------------------------
def gcn_forward(p, cfg, x, graph, pre, dropout_masks=None):
    """layers.py:285-292 + :318-323 (optional nn.LayerNorm over the feature dim, :287-288)."""
    L = cfg.gcn_layer_num
    out = x
    for l in range(L):
        y = torch.bmm(graph, out) @ p[pre + f'gcn_layers.{l}.W.weight'].t() + p[pre + f'gcn_layers.{l}.W.bias']
        if getattr(cfg, 'gcn_layer_norm', False):
            y = F.layer_norm(y, (y.shape[-1],), p[pre + f'gcn_layers.{l}.layer_normalization.weight'],
                             p[pre + f'gcn_layers.{l}.layer_normalization.bias'], 1e-5)
        y = F.relu(y)
        if not cfg.no_gcn_residual:
            y = y + out
        if l < L - 1 and dropout_masks is not None:
            y = y * dropout_masks[f'gcn{l}']
        out = y
    return out


def sue_forward(p, cfg, history_embedding, graph, category_mask, category_indices, candidate,
                pre='user_encoder.', dropout_masks=None):
    """userEncoders.py:68-98 with history_embedding = CNE(history) already computed."""
    B, H, D = history_embedding.shape
    n = candidate.shape[1]
    C1 = cfg.category_num + 1
    category_mask[:, -1] = 1                                                       # :73
    if cfg.user_encoder == 'SUE_wo_GCN':                                           # variantEncoders.py:364-390: no graph
        g = history_embedding
    else:
        proxy = p[pre + 'proxy_node_embedding'].unsqueeze(0).expand(B, -1, -1)
        if dropout_masks is not None:
            proxy = proxy * dropout_masks['proxy']
        x0 = torch.cat([history_embedding, proxy], dim=1)                          # :80
        g = gcn_forward(p, cfg, x0, graph, pre + 'gcn.', dropout_masks) + x0       # :81
        g = g[:, :H, :]                                                            # :82
    if cfg.user_encoder == 'SUE_wo_HCA':                                           # variantEncoders.py:416-419
        u = additive_attention(p, pre + 'attention.', g, None)
        return u.unsqueeze(1).repeat(1, n, 1)
    Au = p[pre + 'intraCluster_K.weight'].shape[0]
    K = g @ p[pre + 'intraCluster_K.weight'].t()                                   # [B,H,Au]   :85
    if (pre + 'intraCluster_K.bias') in p:                                         # SUE_wo_GCN (variantEncoders.py:346)
        K = K + p[pre + 'intraCluster_K.bias']
    Q = candidate @ p[pre + 'intraCluster_Q.weight'].t() + p[pre + 'intraCluster_Q.bias']  # [B,n,Au] :86
    a = torch.einsum('bha,bka->bkh', K, Q) / math.sqrt(float(Au))                  # :87
    idx = category_indices.unsqueeze(1).expand(-1, n, -1)
    alpha = scatter_softmax(a, idx, 2).unsqueeze(3)                                # :88
    intra = scatter_sum(alpha * g.unsqueeze(1), idx, 2, C1)                        # [B,n,C1,D] :89
    f = F.relu(intra @ p[pre + 'clusterFeatureAffine.weight'].t() + p[pre + 'clusterFeatureAffine.bias']) + intra  # :91
    if dropout_masks is not None:
        f = f * dropout_masks['cluster']
    cmask = category_mask.unsqueeze(1).expand(-1, n, -1).reshape(B * n, C1)
    u = scaled_dot_candidate_attention(p, pre + 'interClusterAttention.', f.reshape(B * n, C1, D),
                                       candidate.reshape(B * n, D), cmask)         # :93-97
    return u.view(B, n, D)


# ----------------------------------------------------------------------------------------------
# Model / loss / step (model.py:120-133, trainer.py:64-66,116-120)
# ----------------------------------------------------------------------------------------------
BATCH_FIELDS = ['user_ID', 'user_category', 'user_subCategory', 'user_title_text', 'user_title_mask',
                'user_title_entity', 'user_content_text', 'user_content_mask', 'user_content_entity',
                'user_history_mask', 'user_history_graph', 'user_history_category_mask',
                'user_history_category_indices', 'news_category', 'news_subCategory', 'news_title_text',
                'news_title_mask', 'news_title_entity', 'news_content_text', 'news_content_mask',
                'news_content_entity']


def model_forward(p, cfg, batch, sort_fn=None, lstm_impl='loop', dropout_masks=None):
    """batch: dict over BATCH_FIELDS (CPU tensors; masks are cloned here because the reference
    mutates them in place).  Returns logits [B, n]  (model.py:123-127)."""
    b = {k: (v.clone() if torch.is_tensor(v) and v.dtype == torch.bool else v) for k, v in batch.items()}
    ca = cfg.news_encoder in ('CNE', 'CNE_wo_CS')
    gt = cfg.news_encoder in ('CNE', 'CNE_wo_CA')
    md = cne_modalities(cfg)
    dm = dropout_masks or {}
    news = cne_forward(p, cfg, b['news_title_text'], b['news_title_mask'], b['news_content_text'],
                       b['news_content_mask'], b['news_category'], b['news_subCategory'],
                       sort_fn=sort_fn, lstm_impl=lstm_impl, dropout_masks=dm.get('news'), cross_attention=ca, gate=gt, modalities=md)
    hist = cne_forward(p, cfg, b['user_title_text'], b['user_title_mask'], b['user_content_text'],
                       b['user_content_mask'], b['user_category'], b['user_subCategory'],
                       sort_fn=sort_fn, lstm_impl=lstm_impl, dropout_masks=dm.get('history'), cross_attention=ca, gate=gt, modalities=md)
    user = sue_forward(p, cfg, hist, b['user_history_graph'], b['user_history_category_mask'],
                       b['user_history_category_indices'], news, dropout_masks=dm.get('sue'))
    return (user * news).sum(dim=2)


def loss_fn(logits):
    """trainer.py:64-66."""
    return (-torch.log_softmax(logits, dim=1).select(1, 0)).mean()


def forward_backward(p, cfg, batch, dtype=torch.float32, **kw):
    """Returns (logits, loss, grads dict) with fresh leaf copies of p in ``dtype``."""
    leaf = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    bb = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    logits = model_forward(leaf, cfg, bb, **kw)
    loss = loss_fn(logits)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return logits.detach(), loss.detach(), grads


def clip_and_adam(params, grads, state, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=4.0):
    """trainer.py:118-120: clip_grad_norm_(params, 4) then Adam(lr, wd=0) -- restated directly.
    ``state`` maps name -> (m, v); updated in place; ``step`` is 1-based."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for k, w in params.items():
        g = grads[k] * coef
        m, v = state.setdefault(k, (torch.zeros_like(w), torch.zeros_like(w)))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(m, denom, value=-lr / bc1)
    return total
