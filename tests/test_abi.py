"""CPU tests: the C-ABI library loads and exports every symbol include/nnr_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'nnr_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nnr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from nnr_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(_lib.lib, s), 'missing symbol ' + s
        assert s in _lib.SIGNATURES, 'no ctypes signature for ' + s
    header = open(os.path.join(ROOT, 'include', 'nnr_b200.h')).read()
    assert _lib.lib.nnr_abi_version() == int(re.search(r'#define\s+NNR_ABI_VERSION\s+(\d+)', header).group(1))


def test_graft_entry_build_check_accepts_the_current_library():
    """__graft_entry__.build() asserts the loaded library matches the header's ABI version (no recompilation here)"""
    import importlib
    import subprocess
    g = importlib.import_module('__graft_entry__')
    real = subprocess.run
    try:
        subprocess.run = lambda *a, **k: None           # skip nvcc: the library is already built for this test session
        g.build()
    finally:
        subprocess.run = real


def test_argument_errors_are_reported_without_a_gpu():
    from nnr_b200 import _lib
    rc = _lib.lib.nnr_seq_prepare(None, 0, 0, None, None, None, None)
    assert rc < 0
    assert b'nnr_seq_prepare' in _lib.lib.nnr_last_error()
    a = _lib.GemmArgs()
    assert _lib.lib.nnr_gemm(ctypes.byref(a), None) < 0


def test_plane_producing_entry_points_validate_arguments_without_a_gpu():
    """the entry points that hand their result to nnr_gemm as operand planes (ABI 4): bad arguments are reported, the
    capability query answers from the configuration alone"""
    from nnr_b200 import _lib
    L = _lib.lib
    assert L.nnr_embed_gather_planes_fwd(None, None, None, None, 0, 0, 0, 0, 0, 0.0, 0, 4, None, 0, None) < 0
    assert b'nnr_embed_gather_planes_fwd' in L.nnr_last_error()
    assert L.nnr_lstm_shift_h_planes(None, None, None, None, 0, 0, 0, 0, 4, None, 0, None) < 0
    assert L.nnr_gate_bwd_planes(None, None, None, None, 0, 0, 0, 4, None, 0, None, None, 0, None) < 0
    assert L.nnr_relu_bwd_split_colsum(None, None, 0, 0, 0, 0.0, 0, None, 4, None, 0, None, 0, None, 0, None) < 0
    assert L.nnr_gcn_aggregate_add(None, None, None, None, 0, 0, 0, None, None, None) < 0
    assert L.nnr_lstm_bwd_planes(None, None, None, None, None, None, 0, 0, 200, None, None, None, 0, 4, None, 0, None, None, 0, None) < 0
    assert L.nnr_lstm_bwd_planes_supported(200, 4) in (0, 1)          # 1 unless NNR_LSTM_ALGO=ffma
    assert L.nnr_lstm_bwd_planes_supported(128, 4) == 0               # only hidden_dim 200 is instantiated
    assert L.nnr_lstm_bwd_planes_supported(200, 1) == 0               # the exact-fp32 GEMM has no operand planes
    assert L.nnr_lstm_bwd_planes_workspace_bytes(3520, 200) == 110 * 1600 * 4
    # a GEMM whose operand exists only as planes cannot run on the exact-fp32 kernel
    a = _lib.GemmArgs()
    a.M, a.N, a.K = 64, 64, 64
    a.lda = a.ldb = a.ldc = 64
    a.algo = 1
    a.A_planes = 16
    a.B = 16
    a.C = 16
    assert L.nnr_gemm(ctypes.byref(a), None) < 0
    assert b'planes' in L.nnr_last_error()


def test_ops_reject_cpu_tensors():
    import pytest
    import torch
    from nnr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.rowdot_fwd(torch.zeros(2, 4), torch.zeros(2, 4), 2, 4, torch.zeros(2))
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.nnr.rowdot_fwd(torch.zeros(2, 4), torch.zeros(2, 4), 2, 4, torch.zeros(2))


def test_state_dict_contract():
    import torch
    import nnr_b200
    from oracle import nnr_oracle as O
    cfg = O.make_config(vocabulary_size=300)
    cfg.pretrained_word_embedding = torch.zeros(300, 300)
    m = nnr_b200.Model(cfg)
    m.initialize()
    want = O.alias_state_dict(O.formula_params(cfg))
    sd = m.state_dict()
    assert set(sd) == set(want)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
    assert m.model_name == 'CNE-SUE' and m.news_encoder.auxiliary_loss is None and m.user_encoder.auxiliary_loss is None
    assert m.news_embedding_dim == 900
