"""In-kernel phase stamps of gemm_tc_kernel on the SUE-sized shapes (needs a library built with -DNNR_TC_PROF):
    touch nnr_b200/csrc/gemm_tc.cu && NVCC_EXTRA=-DNNR_TC_PROF bash nnr_b200/csrc/build.sh && python scripts/gemm_phase_prof.py
CTA 0 stamps clock64 at: entry, set-up done, first operand stage landed, MMAs of tile 0 / tile 1 issued, accumulator of tile 0
ready / epilogue of tile 0 done, the same for tile 1, all warps done."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nnr_b200 import ops  # noqa: E402
from nnr_b200._lib import lib  # noqa: E402
from nnr_b200.ops import EPI_BIAS_RELU_RES, EPI_NONE  # noqa: E402

dev = torch.device('cuda:0')
GHZ = 1.965
NAMES = ['entry', 'setup', 'stage0', 'mma_t0', 'mma_t1', 'acc_t0', 'epi_t0', 'acc_t1', 'epi_t1', 'end']
buf = (ctypes.c_ulonglong * 16)()
for (M, N, K, tB, epi) in ((4352, 900, 896, True, EPI_NONE), (4352, 900, 896, True, EPI_BIAS_RELU_RES), (4352, 900, 896, False, EPI_NONE),
                           (4352, 900, 64, True, EPI_NONE), (3520, 400, 384, True, EPI_NONE), (450560, 400, 400, True, EPI_NONE)):
    x = torch.randn(M, K, device=dev)
    W = torch.randn((N, K) if tB else (K, N), device=dev) * 0.05
    out = torch.empty(M, N, device=dev)
    aux = torch.randn(M, N, device=dev)
    aux_out = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    x_pl = ops.tc_split(x, M, K, K)
    w_pl = ops.tc_split(W, W.shape[0], W.shape[1], W.shape[1])
    kw = dict(bias=bias, aux=aux, ldaux=N, aux_out=aux_out, ldaux_out=N, p_drop=0.2, seed=1234) if epi == EPI_BIAS_RELU_RES else {}
    for _ in range(3):
        ops.gemm(x, W, out, M, N, K, K, W.stride(0), N, False, tB, epi, a_planes=x_pl, b_planes=w_pl, **kw)
    torch.cuda.synchronize()
    if not lib.nnr_debug_tc_prof(buf):
        sys.exit('library built without -DNNR_TC_PROF')
    t = [buf[i] for i in range(10)]
    rel = ['%s %.2f' % (NAMES[i], (t[i] - t[0]) / GHZ / 1e3) for i in range(10) if t[i] >= t[0]]
    print('%6dx%4dx%4d %s e%d  us since entry: %s' % (M, N, K, 'NT' if tB else 'NN', epi, '  '.join(rel)), flush=True)
    del x, W, out, aux, aux_out
