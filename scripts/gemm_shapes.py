"""Micro-benchmark of nnr_gemm on the shapes of one CNE+SUE training step (run under ncu for per-kernel times,
or standalone for CUDA-event times)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nnr_b200 import ops

dev = torch.device('cuda:0')
TOK, CAP = 90000, 450560
SHAPES = [  # (M, N, K, transA, transB, m_dev, k_dev)
    (CAP, 1600, 300, False, True, TOK, None),     # input projection fwd
    (CAP, 300, 1600, False, False, TOK, None),    # dE
    (1600, 300, CAP, True, False, None, TOK),     # dW_ih
    (800, 200, CAP, True, False, None, TOK),      # dW_hh
    (CAP, 400, 400, False, True, TOK, None),      # gate fwd
    (CAP, 400, 400, False, False, TOK, None),     # gate dgrad
    (400, 400, CAP, True, False, None, TOK),      # dH
    (CAP, 200, 400, False, True, TOK, None),      # attention projection
    (4352, 900, 900, False, True, None, None),    # GCN W
    (900, 900, 4352, True, False, None, None),    # GCN dW
]
algo = int(os.environ.get('ALGO', '2'))
for (M, N, K, tA, tB, md, kd) in SHAPES:
    A = torch.randn((K, M) if tA else (M, K), device=dev)
    B = torch.randn((N, K) if tB else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    m_dev = torch.tensor([md], dtype=torch.int32, device=dev) if md else None
    k_dev = torch.tensor([kd], dtype=torch.int32, device=dev) if kd else None
    def run():
        ops.gemm(A, B, C, M, N, K, A.stride(0), B.stride(0), N, tA, tB, m_dev=m_dev, k_dev=k_dev, algo=algo)
    run(); run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = 2.0 * (md or M) * N * (kd or K)
    print('%8dx%5dx%8d %s%s  %7.3f ms  %6.1f TFLOP/s' % (M, N, K, 'T' if tA else 'N', 'T' if tB else 'N', ms, fl / ms / 1e9), flush=True)
    del A, B, C
