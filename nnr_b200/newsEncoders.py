"""Drop-in ``NewsEncoder`` / ``CNE`` (reference newsEncoders.py:11-141) on the nnr_b200 CUDA path.

Same constructor signature, sub-module tree (hence ``state_dict`` keys and RNG consumption order),
``initialize()`` distributions, ``forward`` signature, ``news_embedding_dim`` and ``auxiliary_loss``.
``forward`` runs ``engine.CNEFunction``: hand-written sm_100a kernels behind the C ABI, no ATen math
on per-token tensors, no host synchronisation (the reference syncs twice per call for
``lengths.cpu()``, newsEncoders.py:119-120).
"""
import os
import pickle

import torch
import torch.nn as nn

from . import engine
from .layers import Attention, ScaledDotProduct_CandidateAttention


class NewsEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embedding_dim = config.word_embedding_dim
        self.word_embedding = nn.Embedding(num_embeddings=config.vocabulary_size, embedding_dim=self.word_embedding_dim)
        table = getattr(config, 'pretrained_word_embedding', None)    # extension: pass the table directly
        if table is None:
            fn = 'word_embedding-' + str(config.word_threshold) + '-' + str(config.word_embedding_dim) + '-' + config.tokenizer \
                 + '-' + str(config.max_title_length) + '-' + str(config.max_abstract_length) + '-' + config.dataset + '.pkl'
            if os.path.isfile(fn):                                    # reference behaviour: pickle in the cwd
                with open(fn, 'rb') as f:
                    table = pickle.load(f)
        if table is not None:
            self.word_embedding.weight.data.copy_(torch.as_tensor(table))
        self.category_embedding = nn.Embedding(num_embeddings=config.category_num, embedding_dim=config.category_embedding_dim)
        self.subCategory_embedding = nn.Embedding(num_embeddings=config.subCategory_num, embedding_dim=config.subCategory_embedding_dim)
        self.dropout_rate = config.dropout_rate
        self.dropout = nn.Dropout(p=config.dropout_rate, inplace=True)
        self.dropout_ = nn.Dropout(p=config.dropout_rate, inplace=False)
        self.auxiliary_loss = None

    def initialize(self):
        nn.init.uniform_(self.category_embedding.weight, -0.1, 0.1)
        nn.init.uniform_(self.subCategory_embedding.weight, -0.1, 0.1)
        nn.init.zeros_(self.subCategory_embedding.weight[0])

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity, category, subCategory, user_embedding):
        raise Exception('Function forward must be implemented at sub-class')


class CNE(NewsEncoder):
    cross_attention = True
    selective_gate = True
    gate_gain = 'sigmoid'
    modalities = ('title', 'content')

    def __init__(self, config):
        super().__init__(config)
        self.max_title_length = config.max_title_length
        self.max_content_length = config.max_abstract_length
        self.hidden_dim = config.hidden_dim
        self.attention_dim = config.attention_dim
        self.news_embedding_dim = (config.hidden_dim * 2 * len(self.modalities) + config.category_embedding_dim
                                   + config.subCategory_embedding_dim)
        # parameter holders only: the LSTM recurrence runs in nnr_lstm_fwd/bwd, never through cuDNN
        if 'title' in self.modalities:
            self.title_lstm = nn.LSTM(self.word_embedding_dim, self.hidden_dim, batch_first=True, bidirectional=True)
        if 'content' in self.modalities:
            self.content_lstm = nn.LSTM(self.word_embedding_dim, self.hidden_dim, batch_first=True, bidirectional=True)
        if self.selective_gate:
            self.title_H = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=False)
            self.title_M = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=True)
            self.content_H = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=False)
            self.content_M = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=True)
        if 'title' in self.modalities:
            self.title_self_attention = Attention(self.hidden_dim * 2, config.attention_dim)
        if 'content' in self.modalities:
            self.content_self_attention = Attention(self.hidden_dim * 2, config.attention_dim)
        if self.cross_attention:
            self.title_cross_attention = ScaledDotProduct_CandidateAttention(self.hidden_dim * 2, self.hidden_dim * 2, config.attention_dim)
            self.content_cross_attention = ScaledDotProduct_CandidateAttention(self.hidden_dim * 2, self.hidden_dim * 2, config.attention_dim)

    def initialize(self):
        super().initialize()
        for lstm in [getattr(self, x + '_lstm') for x in self.modalities]:
            for parameter in lstm.parameters():
                if len(parameter.size()) >= 2:
                    nn.init.orthogonal_(parameter.data)
                else:
                    nn.init.zeros_(parameter.data)
        if self.selective_gate:
            gain = nn.init.calculate_gain(self.gate_gain) if self.gate_gain else 1.0
            nn.init.xavier_uniform_(self.title_H.weight, gain=gain)
            nn.init.xavier_uniform_(self.title_M.weight, gain=gain)
            nn.init.zeros_(self.title_M.bias)
            nn.init.xavier_uniform_(self.content_H.weight, gain=gain)
            nn.init.xavier_uniform_(self.content_M.weight, gain=gain)
            nn.init.zeros_(self.content_M.bias)
        for x in self.modalities:
            getattr(self, x + '_self_attention').initialize()
        if self.cross_attention:
            self.title_cross_attention.initialize()
            self.content_cross_attention.initialize()

    def _params(self):
        # the Parameter OBJECTS are fixed for the life of the module (optimizers / TrainStep rebind .data, not the
        # objects), so the module tree is walked once, not on every forward
        cached = self.__dict__.get('_param_list')
        if cached is None:
            names = engine.cne_param_names(self.cross_attention, self.selective_gate, self.modalities)
            sd = dict(self.named_parameters())
            cached = [sd[k] for k in names]
            self.__dict__['_param_list'] = cached
        return cached

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity, category, subCategory, user_embedding):
        return self.encode_calls([(title_text, title_mask, content_text, content_mask, category, subCategory)])[0]

    def encode_calls(self, calls):
        """Encode several reference-style ``news_encoder(...)`` calls with ONE kernel schedule.

        Each element of ``calls`` is (title_text [B,n,T], title_mask, content_text [B,n,A], content_mask, category
        [B,n], subCategory [B,n]) and is its own pairing domain for the cross-selective gate (newsEncoders.py:112-115,
        124-129), exactly as if ``forward`` had been called once per element; everything else is row-independent,
        so candidates (B*5 rows) and history (B*50 rows) share the GEMM / LSTM launches.  Returns one
        [B, n, news_embedding_dim] tensor per call.  Masks are modified in place like the reference (:108-109).
        """
        if not calls[0][0].is_cuda:
            raise RuntimeError('nnr_b200.CNE runs on CUDA (sm_100a) only; there is no CPU path')
        i32 = torch.int32
        T, A = self.max_title_length, self.max_content_length
        domains, shapes, start = [], [], 0
        for c in calls:
            B, n = c[0].size(0), c[0].size(1)
            domains.append((start, B * n))
            shapes.append((B, n))
            start += B * n
        if len(calls) == 1:
            tt, tm, ct, cm, cat, sub = calls[0]
            tt, ct, cat, sub = tt.to(i32).contiguous(), ct.to(i32).contiguous(), cat.to(i32), sub.to(i32)
        else:
            for c in calls:                                    # the caller's masks are mutated like the reference does
                c[1][..., 0] = True
                c[3][..., 0] = True
            tt = torch.cat([c[0].reshape(-1, T).to(i32) for c in calls])
            tm = torch.cat([c[1].reshape(-1, T) for c in calls])
            ct = torch.cat([c[2].reshape(-1, A).to(i32) for c in calls])
            cm = torch.cat([c[3].reshape(-1, A) for c in calls])
            cat = torch.cat([c[4].reshape(-1).to(i32) for c in calls])
            sub = torch.cat([c[5].reshape(-1).to(i32) for c in calls])
        meta = dict(N=start, T=T, A_len=A, E=self.word_embedding_dim, Hd=self.hidden_dim, att=self.attention_dim,
                    training=self.training, p_drop=float(self.dropout_rate), cross_attention=self.cross_attention,
                    gate=self.selective_gate, modalities=self.modalities, domains=domains)
        rep = engine.CNEFunction.apply(meta, tt, tm, ct, cm, cat, sub, *self._params())
        return [rep[s:s + cnt].view(B, n, self.news_embedding_dim) for (s, cnt), (B, n) in zip(domains, shapes)]
