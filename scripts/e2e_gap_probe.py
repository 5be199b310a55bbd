"""Where the gap between the device-resident step and the end-to-end step comes from: the same captured step timed
(a) on device-resident packed batches, (b) + every step's loss read back through LossLog, (c) + the pinned host batch prefetched
on the copy stream (= bench.py's e2e), (d) = (c) with a blocking float(loss)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nnr_b200
from nnr_b200.synthetic import SyntheticMIND, batch_args
from nnr_b200.trainer import LossLog, PackedBatch, TrainStep

sys.argv = [sys.argv[0]]
a = bench.parse()
dev = torch.device('cuda:0')
cfg = bench.make_config(a)
syn = SyntheticMIND(news_num=20000, vocabulary_size=a.vocab, lengths=a.lengths, seed=0)
cfg.pretrained_word_embedding = syn.word_table()
torch.manual_seed(1234)
model = nnr_b200.Model(cfg); model.initialize(); model.to(dev)
ts = TrainStep(model, lr=1e-4, gradient_clip_norm=4.0, cuda_graph=True)
host = [syn.batch(a.batch, seed=i) for i in range(4)]
devp = [PackedBatch.pack(batch_args(b, dev), device=dev) for b in host]
hostp = [PackedBatch.pack(batch_args(b), pin=True) for b in host]
for i in range(4):
    ts.step(devp[i % 4])
torch.cuda.synchronize()
K = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, label):
    torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print('%-58s %.3f ms/step' % (label, e0.elapsed_time(e1) / K), flush=True)


def a_():
    for i in range(K):
        ts.step(devp[i % 4])


def b_():
    log = LossLog()
    for i in range(K):
        log.push(ts.step(devp[i % 4]))
    log.drain()


def c_(blocking=False):
    log = LossLog()
    nxt = ts.prefetch(hostp[0])
    for i in range(K):
        loss = ts.step(nxt)
        if not blocking:
            log.push(loss)
        if i + 1 < K:
            nxt = ts.prefetch(hostp[(i + 1) % 4])
        if blocking:
            loss.item()
    log.drain()


for rep in range(2):
    timed(a_, '(a) device-resident packed batches')
    timed(b_, '(b) + loss of every step through LossLog')
    timed(c_, '(c) + pinned host batch prefetched on the copy stream')
    timed(lambda: c_(True), '(d) (c) with a blocking float(loss) per step')
