"""The K = 400 token-level GEMMs of the CNE step in isolation (selective gate forward `e4`, its dgrad with the add-aux
epilogue `e5`, the attention affine `e2`): CUDA-event time per call, and a convenient target for
`ncu --set full --import-source on -k regex:gemm_tc_kernel`.

    python scripts/gemm_gate_probe.py [tokens] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nnr_b200 import ops  # noqa: E402
from nnr_b200.ops import EPI_ADD_AUX, EPI_BIAS_TANH, EPI_GATE, EPI_NONE  # noqa: E402

dev = torch.device('cuda:0')
TOK = int(sys.argv[1]) if len(sys.argv) > 1 else 84000
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N_NEWS, L, D2, A = 3520, 128, 400, 200
CAP = N_NEWS * L
torch.manual_seed(0)
h = torch.randn(CAP, D2, device=dev)
W = torch.randn(D2, D2, device=dev) * 0.05
W1 = torch.randn(A, D2, device=dev) * 0.05
mproj = torch.randn(N_NEWS, D2, device=dev)
tok_row = (torch.arange(CAP, device=dev) // (TOK // N_NEWS + 1)).clamp_(max=N_NEWS - 1).to(torch.int32)
ntok = torch.tensor([TOK], dtype=torch.int32, device=dev)
g = torch.empty(CAP, D2, device=dev)
hg = torch.empty(CAP, D2, device=dev)
u = torch.empty(CAP, A, device=dev)
b1 = torch.randn(A, device=dev)
h_pl = ops.tc_split(h, CAP, D2, D2, ntok)
w_pl = ops.tc_split(W, D2, D2, D2)
w1_pl = ops.tc_split(W1, A, D2, D2)


def gate():
    ops.gemm(h, W, hg, CAP, D2, D2, D2, D2, D2, False, True, EPI_GATE, rowbias=mproj, ldrowbias=D2, rowmap=tok_row, aux=h,
             ldaux=D2, aux_out=g, ldaux_out=D2, m_dev=ntok, a_planes=h_pl, b_planes=w_pl)


def dgrad_add():
    ops.gemm(h, W, hg, CAP, D2, D2, D2, D2, D2, False, False, EPI_ADD_AUX, aux=g, ldaux=D2, m_dev=ntok, a_planes=h_pl,
             b_planes=w_pl)


def plain():
    ops.gemm(h, W, hg, CAP, D2, D2, D2, D2, D2, False, True, EPI_NONE, m_dev=ntok, a_planes=h_pl, b_planes=w_pl)


def affine_tanh():
    ops.gemm(h, W1, u, CAP, A, D2, D2, D2, A, False, True, EPI_BIAS_TANH, bias=b1, m_dev=ntok, a_planes=h_pl, b_planes=w1_pl)


for name, fn, n_out in (('gate e4 NT', gate, D2), ('dgrad+aux e5 NN', dgrad_add, D2), ('plain e0 NT', plain, D2),
                        ('affine tanh e2 NT', affine_tanh, A)):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    print('%-20s tokens %6d: %7.3f ms  %6.1f TFLOP/s algorithmic' % (name, TOK, ms, 2.0 * TOK * n_out * D2 / ms / 1e9), flush=True)
