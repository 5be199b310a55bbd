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

# Cases whose inputs are NOT stored (they are regenerated from the seeded synthetic corpus and checked against a stored
# SHA-256): name -> (config overrides, corpus kwargs, batch size, candidates per impression or None, batch seed, samples per
# gradient digest).  `config2` is BASELINE.json configs[1] (batch 64, V = 40 000, MIND-like lengths, gcn 4) at dropout 0.
BIG_CASES = {
    'config2': (dict(vocabulary_size=40000, subCategory_num=285, gcn_layer_num=4), dict(news_num=20000, lengths='mind', seed=0), 64, None, 21, 64),
    'full_b8': (dict(vocabulary_size=5000, subCategory_num=285, gcn_layer_num=4), dict(news_num=2000, lengths='full', seed=23), 8, None, 22, 64),
}
# small cases added in round 2 (inputs stored like the round-1 cases)
GOLDEN_CASES.update({
    'gcn5': dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24, subCategory_num=30, gcn_layer_num=5),
    'gcn7': dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24, subCategory_num=30, gcn_layer_num=7),
    'no_residual': dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24, subCategory_num=30,
                        gcn_layer_num=3, no_gcn_residual=True),
    'layer_norm': dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24, subCategory_num=30,
                       gcn_layer_num=3, gcn_layer_norm=True),
})


def sample_positions(numel, n=SAMPLE):
    return (np.arange(n, dtype=np.int64) * 2654435761 + 12345) % max(numel, 1)


def corpus_for(cfg, kw):
    from nnr_b200.synthetic import SyntheticMIND
    return SyntheticMIND(vocabulary_size=cfg.vocabulary_size, category_num=cfg.category_num,
                         subCategory_num=cfg.subCategory_num, max_title_length=cfg.max_title_length,
                         max_abstract_length=cfg.max_abstract_length, max_history_num=cfg.max_history_num,
                         negative_sample_num=cfg.negative_sample_num, **kw)


def batch_sha256(batch):
    import hashlib
    h = hashlib.sha256()
    for k in O.BATCH_FIELDS:
        v = batch.get(k)
        if torch.is_tensor(v):
            h.update(k.encode())
            h.update(np.ascontiguousarray(v.numpy()).tobytes())
    return h.hexdigest()


def load_big_golden(name):
    """(cfg, batch regenerated from the seeds and verified against the stored hash, npz)"""
    over, ckw, B, n, bseed, _ = BIG_CASES[name]
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = O.make_config(**over)
    batch = corpus_for(cfg, ckw).batch(B, news_num=n, seed=bseed)
    assert batch_sha256(batch) == str(z['in_sha256']), 'synthetic generator drifted: %s inputs differ from the golden run' % name
    return cfg, batch, z


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = O.make_config(**GOLDEN_CASES[name])
    batch = {}
    for k in O.BATCH_FIELDS:
        batch[k] = torch.from_numpy(z['in_' + k]) if ('in_' + k) in z.files else None
    return cfg, batch, z


def grad_digest(g, n=SAMPLE):
    flat = g.detach().reshape(-1).double().cpu()
    pos = sample_positions(flat.numel(), n)
    return np.concatenate([[flat.sum().item(), flat.abs().sum().item(), flat.abs().max().item()],
                           flat[torch.from_numpy(pos)].numpy()])


def rel_err(a, b):
    """max |a-b| / max |b|  (SURVEY 8c metric)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
