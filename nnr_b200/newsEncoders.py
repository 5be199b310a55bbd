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
    gate_gain = 'sigmoid'

    def __init__(self, config):
        super().__init__(config)
        self.max_title_length = config.max_title_length
        self.max_content_length = config.max_abstract_length
        self.hidden_dim = config.hidden_dim
        self.attention_dim = config.attention_dim
        self.news_embedding_dim = config.hidden_dim * 4 + config.category_embedding_dim + config.subCategory_embedding_dim
        # parameter holders only: the LSTM recurrence runs in nnr_lstm_fwd/bwd, never through cuDNN
        self.title_lstm = nn.LSTM(self.word_embedding_dim, self.hidden_dim, batch_first=True, bidirectional=True)
        self.content_lstm = nn.LSTM(self.word_embedding_dim, self.hidden_dim, batch_first=True, bidirectional=True)
        self.title_H = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=False)
        self.title_M = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=True)
        self.content_H = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=False)
        self.content_M = nn.Linear(self.hidden_dim * 2, self.hidden_dim * 2, bias=True)
        self.title_self_attention = Attention(self.hidden_dim * 2, config.attention_dim)
        self.content_self_attention = Attention(self.hidden_dim * 2, config.attention_dim)
        if self.cross_attention:
            self.title_cross_attention = ScaledDotProduct_CandidateAttention(self.hidden_dim * 2, self.hidden_dim * 2, config.attention_dim)
            self.content_cross_attention = ScaledDotProduct_CandidateAttention(self.hidden_dim * 2, self.hidden_dim * 2, config.attention_dim)

    def initialize(self):
        super().initialize()
        for lstm in (self.title_lstm, self.content_lstm):
            for parameter in lstm.parameters():
                if len(parameter.size()) >= 2:
                    nn.init.orthogonal_(parameter.data)
                else:
                    nn.init.zeros_(parameter.data)
        gain = nn.init.calculate_gain(self.gate_gain) if self.gate_gain else 1.0
        nn.init.xavier_uniform_(self.title_H.weight, gain=gain)
        nn.init.xavier_uniform_(self.title_M.weight, gain=gain)
        nn.init.zeros_(self.title_M.bias)
        nn.init.xavier_uniform_(self.content_H.weight, gain=gain)
        nn.init.xavier_uniform_(self.content_M.weight, gain=gain)
        nn.init.zeros_(self.content_M.bias)
        self.title_self_attention.initialize()
        self.content_self_attention.initialize()
        if self.cross_attention:
            self.title_cross_attention.initialize()
            self.content_cross_attention.initialize()

    def _params(self):
        names = engine.CNE_PARAM_NAMES + (engine.CNE_CROSS_PARAM_NAMES if self.cross_attention else [])
        sd = dict(self.named_parameters())
        return [sd[k] for k in names]

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity, category, subCategory, user_embedding):
        if not title_text.is_cuda:
            raise RuntimeError('nnr_b200.CNE runs on CUDA (sm_100a) only; there is no CPU path')
        B, n = title_text.size(0), title_text.size(1)
        meta = dict(N=B * n, T=self.max_title_length, A_len=self.max_content_length, E=self.word_embedding_dim,
                    Hd=self.hidden_dim, att=self.attention_dim, training=self.training, p_drop=float(self.dropout_rate),
                    cross_attention=self.cross_attention)
        i32 = torch.int32
        rep = engine.CNEFunction.apply(
            meta, title_text.to(i32).contiguous(), title_mask, content_text.to(i32).contiguous(), content_mask,
            category.to(i32), subCategory.to(i32), *self._params())
        return rep.view(B, n, self.news_embedding_dim)
