"""In-kernel phase timing of the tensor-core LSTM kernels (needs a library built with -DNNR_LSTM_PROF):
    NVCC_EXTRA=-DNNR_LSTM_PROF bash nnr_b200/csrc/build.sh && python scripts/lstm_phase_prof.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nnr_b200 import ops
from nnr_b200._lib import lib

dev = torch.device('cuda:0')
Hd = 200
for (N, L) in [(32, 64)]:
    lens = torch.full((N,), L, dtype=torch.int32, device=dev)
    off = (torch.arange(N + 1, device=dev) * L).to(torch.int32)
    order = torch.arange(N, dtype=torch.int32, device=dev)
    gx = torch.randn(N * L, 8 * Hd, device=dev) * 0.1
    w_hh = torch.randn(2, 4 * Hd, Hd, device=dev) * 0.05
    h = torch.empty(N * L, 2 * Hd, device=dev); cst = torch.empty(N * L, 2 * Hd, device=dev); cn = torch.empty(N, 2 * Hd, device=dev)
    dh = torch.randn(N * L, 2 * Hd, device=dev) * 0.1; dcn = torch.randn(N, 2 * Hd, device=dev) * 0.1
    for _ in range(2):
        ops.lstm_fwd(gx, w_hh, lens, off, order, N, L, Hd, h, cst, cn)
        ops.lstm_bwd(gx, cst, w_hh, lens, off, order, N, L, Hd, dh, dcn)
    out = (ctypes.c_ulonglong * 16)()
    lib.nnr_debug_lstm_prof(out)
    v = list(out)
    print('N=%d L=%d' % (N, L))
    print('  fwd cycles: wait %d  mma %d  cell+stores %d  stage+sync+send %d  gx-prefetch %d  loop %d' % tuple(v[0:6]))
    print('  bwd cycles: c-loads %d  wait-rfull+cp %d  phase1 %d  sync+storeout+prefetch %d  mma %d  loop+bulk-issue %d  rfree-wait %d  stage+sync %d' % tuple(v[8:16]))
