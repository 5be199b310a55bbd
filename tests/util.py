"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from oracle import nnr_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

GOLDEN_CASES = {
    'tiny': dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24,
                 subCategory_num=30, gcn_layer_num=3),
    'mind_shape': dict(vocabulary_size=2000, subCategory_num=285, gcn_layer_num=4),
    'dev_shape': dict(vocabulary_size=800, max_history_num=10, max_title_length=16, max_abstract_length=40,
                      subCategory_num=40, gcn_layer_num=2),
    'ablation': dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                     subCategory_num=30, gcn_layer_num=1, news_encoder='CNE_wo_CA', user_encoder='SUE_wo_HCA'),
    'wo_cs': dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                  subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_wo_CS'),
    'wo_gcn': dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                   subCategory_num=30, gcn_layer_num=2, user_encoder='SUE_wo_GCN'),
    'title_only': dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                       subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_Title'),
    'content_only': dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                         subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_Content'),
}
SAMPLE = 8


def sample_positions(numel):
    return (np.arange(SAMPLE, dtype=np.int64) * 2654435761 + 12345) % max(numel, 1)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = O.make_config(**GOLDEN_CASES[name])
    batch = {}
    for k in O.BATCH_FIELDS:
        batch[k] = torch.from_numpy(z['in_' + k]) if ('in_' + k) in z.files else None
    return cfg, batch, z


def grad_digest(g):
    flat = g.detach().reshape(-1).double().cpu()
    pos = sample_positions(flat.numel())
    return np.concatenate([[flat.sum().item(), flat.abs().sum().item(), flat.abs().max().item()],
                           flat[torch.from_numpy(pos)].numpy()])


def rel_err(a, b):
    """max |a-b| / max |b|  (SURVEY 8c metric)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
