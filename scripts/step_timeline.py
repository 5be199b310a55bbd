"""Kernel timeline of ONE replay of the captured training step (torch.profiler / CUPTI activity records: start, duration and
stream of every kernel inside the graph launch), written as CSV.  Unlike the ncu launch list the kernels run warm, back to
back and -- with the two-lane schedule (engine.Lanes) -- concurrently, so this is where overlap and idle gaps are read from.

    python scripts/step_timeline.py [out.csv] [bench.py flags]
"""
import csv
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import nnr_b200
from nnr_b200.synthetic import SyntheticMIND, batch_args
from nnr_b200.trainer import PackedBatch, TrainStep

out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'gpurun_out/step_timeline.csv'
sys.argv = [sys.argv[0]] + [x for x in sys.argv[1:] if x != out]
a = bench.parse()
dev = torch.device('cuda:0')
cfg = bench.make_config(a)
syn = SyntheticMIND(news_num=20000, vocabulary_size=a.vocab, lengths=a.lengths, seed=0)
cfg.pretrained_word_embedding = syn.word_table()
torch.manual_seed(1234)
if a.dtype == 'bf16':
    os.environ['NNR_GEMM_ALGO'] = 'bf16'
model = nnr_b200.Model(cfg); model.initialize(); model.to(dev)
ts = TrainStep(model, lr=1e-4, gradient_clip_norm=4.0, cuda_graph=True)
devp = [PackedBatch.pack(batch_args(syn.batch(a.batch, seed=i), dev), device=dev) for i in range(4)]
for i in range(6):
    ts.step(devp[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        ts.step(devp[i % 4])
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range is not None:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, 'device_index', 0), ev.name))
rows.sort()
# keep the LAST replay: kernels after the last adam_kernel but one
adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[3]]
lo = adam[-2] + 1 if len(adam) >= 2 else 0
hi = adam[-1] + 1
# the weight-plane refresh follows the optimizer kernel inside the same graph: include everything up to the next memcpy gap
rows = rows[lo:hi + 2]
t0 = rows[0][0]
os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['start_us', 'dur_us', 'name'])
    for s, d, _, n in rows:
        w.writerow(['%.2f' % (s - t0), '%.2f' % d, n[:160]])
busy = sum(r[1] for r in rows)
span = rows[-1][0] + rows[-1][1] - t0
print('kernels %d  span %.1f us  sum of durations %.1f us' % (len(rows), span, busy))
