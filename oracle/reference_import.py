"""TEST INFRASTRUCTURE -- import the UNMODIFIED reference modules from /root/reference.

Only usable where the reference checkout exists (the build container).  Nothing on the GPU box
may call this at run time; it exists to (a) validate ``nnr_oracle`` against the real thing and
(b) generate ``tests/golden/*`` (``tests/golden/make_golden.py``).

Shims:
  * ``torch_scatter`` (pinned 2.0.9, not installed): a module object exposing ``scatter_sum`` and
    ``scatter_softmax`` built on ``oracle.nnr_oracle``'s restatement of the published semantics.
  * ``config.Config`` cannot be constructed (downloads data, asserts a GPU): a plain namespace
    with the attributes the hot path reads is passed instead.
  * ``NewsEncoder.__init__`` unpickles the word table from the cwd (newsEncoders.py:16-17): a
    synthetic table is written to a temp dir and the cwd is switched for the constructor call.
"""
import contextlib
import os
import pickle
import sys
import tempfile
import types

import torch

REFERENCE_ROOT = os.environ.get('NNR_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'newsEncoders.py'))


def _install_torch_scatter_shim():
    if 'torch_scatter' in sys.modules:
        return
    from oracle import nnr_oracle as O
    mod = types.ModuleType('torch_scatter')

    def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
        return O.scatter_sum(src, index, dim, dim_size)

    def scatter_softmax(src, index, dim=-1, dim_size=None):
        return O.scatter_softmax(src, index, dim)

    mod.scatter_sum = scatter_sum
    mod.scatter_softmax = scatter_softmax
    sys.modules['torch_scatter'] = mod


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def import_reference():
    """Returns the reference's ``model`` module (with newsEncoders/userEncoders/layers loaded)."""
    if not available():
        raise RuntimeError('reference checkout not present at ' + REFERENCE_ROOT)
    _install_torch_scatter_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    argv = sys.argv
    sys.argv = argv[:1]            # config.py builds an argparse parser at construction only; be safe
    try:
        import model as ref_model  # noqa: E402  (the reference's model.py)
    finally:
        sys.argv = argv
    return ref_model


def build_reference_model(cfg, params=None, word_table=None):
    """Model(cfg) from the reference; optionally load ``params`` (unique-key dict, see
    nnr_oracle.param_shapes) into it.  Returns the nn.Module on CPU."""
    from oracle import nnr_oracle as O
    ref_model = import_reference()
    if word_table is None:
        word_table = (params['news_encoder.word_embedding.weight'] if params is not None
                      else torch.zeros(cfg.vocabulary_size, cfg.word_embedding_dim))
    with tempfile.TemporaryDirectory() as d:
        fn = 'word_embedding-%s-%s-%s-%s-%s-%s.pkl' % (cfg.word_threshold, cfg.word_embedding_dim, cfg.tokenizer,
                                                     cfg.max_title_length, cfg.max_abstract_length, cfg.dataset)
        with open(os.path.join(d, fn), 'wb') as f:
            pickle.dump(word_table.detach().float().clone(), f)
        with _cwd(d):
            m = ref_model.Model(cfg)
    m.initialize()
    if params is not None:
        missing, unexpected = m.load_state_dict(O.alias_state_dict(params), strict=True)
    return m


def run_reference(m, batch, sort_fn=None):
    """Call the reference model with the 21 positional tensors (masks cloned: it mutates them).
    ``sort_fn`` optionally replaces ``torch.sort`` for the duration of the call (used to fix the
    tie-breaking of the length sort, SURVEY.md finding 2)."""
    from oracle import nnr_oracle as O
    args = []
    for k in O.BATCH_FIELDS:
        v = batch.get(k)
        args.append(v.clone() if torch.is_tensor(v) else v)
    if sort_fn is None:
        return m(*args)
    orig = torch.sort
    torch.sort = sort_fn
    try:
        return m(*args)
    finally:
        torch.sort = orig
