"""Generate tests/golden/*.npz by running the UNMODIFIED reference (needs /root/reference).

    python tests/golden/make_golden.py

Each case stores the integer inputs, the reference's eval-mode logits (default torch.sort and
stable sort injected), and for train mode at dropout_rate = 0 the loss and a digest of every
parameter gradient (sum, abs-sum, max-abs and 8 sampled entries).  Weights are NOT stored: they
come from ``oracle.nnr_oracle.formula_params`` (a closed formula of the element index), so the
fixtures stay small and do not depend on torch's RNG stream.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nnr_oracle as O            # noqa: E402
from oracle import reference_import as R      # noqa: E402
from nnr_b200.synthetic import SyntheticMIND  # noqa: E402
from tests import util as U                   # noqa: E402

CASES = {
    # name: (config overrides, corpus kwargs, batch size, news_num, seeds)
    'tiny': (dict(vocabulary_size=500, max_history_num=6, max_title_length=12, max_abstract_length=24,
                  subCategory_num=30, gcn_layer_num=3),
             dict(news_num=200, lengths='uniform', seed=3), 3, None, 1),
    'mind_shape': (dict(vocabulary_size=2000, subCategory_num=285, gcn_layer_num=4),
                   dict(news_num=600, lengths='mind', seed=5), 2, None, 2),
    'dev_shape': (dict(vocabulary_size=800, max_history_num=10, max_title_length=16, max_abstract_length=40,
                       subCategory_num=40, gcn_layer_num=2),
                  dict(news_num=300, lengths='uniform', seed=7), 4, 1, 3),
    'ablation': (dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                      subCategory_num=30, gcn_layer_num=1, news_encoder='CNE_wo_CA', user_encoder='SUE_wo_HCA'),
                 dict(news_num=200, lengths='uniform', seed=9), 3, None, 4),
    'wo_cs': (dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                   subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_wo_CS'),
              dict(news_num=200, lengths='mind', seed=11), 3, None, 5),
    'wo_gcn': (dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                    subCategory_num=30, gcn_layer_num=2, user_encoder='SUE_wo_GCN'),
               dict(news_num=200, lengths='mind', seed=13), 3, None, 6),
    'title_only': (dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                        subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_Title'),
                   dict(news_num=200, lengths='mind', seed=15), 3, None, 7),
    'content_only': (dict(vocabulary_size=500, max_history_num=8, max_title_length=10, max_abstract_length=20,
                          subCategory_num=30, gcn_layer_num=2, news_encoder='CNE_Content'),
                     dict(news_num=200, lengths='mind', seed=17), 3, None, 8),
}

CASES.update({
    'gcn5': (U.GOLDEN_CASES['gcn5'], dict(news_num=200, lengths='uniform', seed=31), 3, None, 9),
    'gcn7': (U.GOLDEN_CASES['gcn7'], dict(news_num=200, lengths='uniform', seed=33), 3, None, 10),
    'no_residual': (U.GOLDEN_CASES['no_residual'], dict(news_num=200, lengths='uniform', seed=35), 3, None, 11),
    'layer_norm': (U.GOLDEN_CASES['layer_norm'], dict(news_num=200, lengths='uniform', seed=37), 3, None, 12),
})

SAMPLE = 8


def sample_positions(numel):
    return (np.arange(SAMPLE, dtype=np.int64) * 2654435761 + 12345) % max(numel, 1)


def corpus_for(cfg, kw):
    return SyntheticMIND(vocabulary_size=cfg.vocabulary_size, category_num=cfg.category_num,
                         subCategory_num=cfg.subCategory_num, max_title_length=cfg.max_title_length,
                         max_abstract_length=cfg.max_abstract_length, max_history_num=cfg.max_history_num,
                         negative_sample_num=cfg.negative_sample_num, **kw)


def make_case(name):
    over, ckw, B, n, bseed = CASES[name]
    cfg = O.make_config(**over)
    syn = corpus_for(cfg, ckw)
    batch = syn.batch(B, news_num=n, seed=bseed)
    p = O.formula_params(cfg)
    out = {}
    for k in O.BATCH_FIELDS + ['history_len']:
        v = batch.get(k)
        if torch.is_tensor(v):
            out['in_' + k] = v.numpy()
    m = R.build_reference_model(cfg, p)
    m.eval()
    with torch.no_grad():
        out['logits_default_sort'] = R.run_reference(m, batch).numpy()
        out['logits_stable_sort'] = R.run_reference(m, batch, sort_fn=O.stable_sort).numpy()
    cfg.dropout_rate = 0.0
    m = R.build_reference_model(cfg, p)
    m.train()
    logits = R.run_reference(m, batch, sort_fn=O.stable_sort)
    loss = O.loss_fn(logits)
    loss.backward()
    out['train_logits'] = logits.detach().numpy()
    out['train_loss'] = loss.detach().numpy()
    named = dict(m.named_parameters())
    for k in O.param_shapes(cfg):
        g = named[k].grad
        g = torch.zeros_like(named[k]) if g is None else g
        flat = g.reshape(-1).double()
        pos = sample_positions(flat.numel())
        out['grad_' + k] = np.concatenate([[flat.sum().item(), flat.abs().sum().item(), flat.abs().max().item()],
                                           flat[torch.from_numpy(pos)].numpy()])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz'), **out)
    print(name, 'loss', float(loss), 'logits', out['logits_stable_sort'].reshape(-1)[:4])


def make_big_case(name):
    """BASELINE-shaped cases: inputs regenerated from seeds (only their SHA-256 is stored); eval logits, and at dropout 0 the
    train-mode logits, loss and a digest (sum, abs-sum, max-abs, S sampled entries) of every parameter gradient of the
    UNMODIFIED reference in fp32 AND with the reference module cast to fp64 (the tie-break judge of SURVEY 8c)"""
    over, ckw, B, n, bseed, S = U.BIG_CASES[name]
    cfg = O.make_config(**over)
    batch = U.corpus_for(cfg, ckw).batch(B, news_num=n, seed=bseed)
    p = O.formula_params(cfg)
    out = {'in_sha256': np.array(U.batch_sha256(batch))}
    torch.set_num_threads(os.cpu_count())
    m = R.build_reference_model(cfg, p)
    m.eval()
    with torch.no_grad():
        out['logits_stable_sort'] = R.run_reference(m, batch, sort_fn=O.stable_sort).numpy()
    cfg.dropout_rate = 0.0
    for tag, dtype in (('', torch.float32), ('64', torch.float64)):
        m = R.build_reference_model(cfg, p)
        m.to(dtype)
        m.train()
        b = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
        logits = R.run_reference(m, b, sort_fn=O.stable_sort)
        loss = O.loss_fn(logits)
        loss.backward()
        out['train_logits' + tag] = logits.detach().numpy()
        out['train_loss' + tag] = loss.detach().numpy()
        named = dict(m.named_parameters())
        for k in O.param_shapes(cfg):
            g = named[k].grad
            g = torch.zeros_like(named[k]) if g is None else g
            out['grad%s_%s' % (tag, k)] = U.grad_digest(g, S)
        print(name, 'fp' + (tag or '32'), 'loss', float(loss), flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz'), **out)


def make_pins():
    """graph_flags.npz / metrics.npz: outputs of the reference's own source lines (MIND_corpus.py:179-213 executed under
    stub variables, evaluate.py imported live, util.py:57-60 executed) on the seeded cases of tests/test_pins.py"""
    from tests import test_pins as T
    here = os.path.dirname(os.path.abspath(__file__))
    code = compile(T.reference_graph_source(), 'MIND_corpus.py:179-213', 'exec')
    cases, H, C = T.graph_cases()
    out = {}
    for fi, flags in enumerate(T.FLAG_SETS):
        g, m, i = zip(*[T.run_reference_graph(code, cats, H, C, flags) for cats in cases])
        out['graph_%d' % fi], out['mask_%d' % fi], out['idx_%d' % fi] = np.stack(g), np.stack(m), np.stack(i)
    np.savez_compressed(os.path.join(here, 'graph_flags.npz'), **out)
    labels, scores = T.metric_cases()
    ref, ranks = T.reference_metrics(labels, scores)
    np.savez_compressed(os.path.join(here, 'metrics.npz'), metrics=np.array([float(x) for x in ref], dtype=np.float64),
                        ranks=np.concatenate([np.array(r) for r in ranks]))
    print('pins', [float(x) for x in ref])


if __name__ == '__main__':
    torch.manual_seed(0)
    names = sys.argv[1:] or (list(CASES) + ['pins'] + list(U.BIG_CASES))
    for name in names:
        if name == 'pins':
            make_pins()
        elif name in U.BIG_CASES:
            make_big_case(name)
        else:
            make_case(name)
