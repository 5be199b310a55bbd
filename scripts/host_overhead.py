"""Host-side enqueue time of one training step (no synchronisation inside the loop) against the device time:
if the two are close, the step is launch-bound and kernel speed-ups stop showing up in throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nnr_b200
from nnr_b200.synthetic import SyntheticMIND, batch_args
from nnr_b200.trainer import TrainStep

sys.argv = [sys.argv[0]]
a = bench.parse()
dev = torch.device('cuda:0')
cfg = bench.make_config(a)
syn = SyntheticMIND(news_num=20000, vocabulary_size=a.vocab, lengths=a.lengths, seed=0)
cfg.pretrained_word_embedding = syn.word_table()
model = nnr_b200.Model(cfg); model.initialize(); model.to(dev)
ts = TrainStep(model, lr=1e-4, gradient_clip_norm=4.0)
devb = [batch_args(syn.batch(a.batch, seed=i), dev) for i in range(4)]
for i in range(4):
    ts.step(*devb[i % 4])
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(10):
        ts.step(*devb[i % 4])
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('host enqueue %.2f ms/step   device %.2f ms/step   wall %.2f ms/step' % ((t1 - t0) * 100, e0.elapsed_time(e1) / 10, (t2 - t0) * 100), flush=True)
# forward / backward / optimizer split of the host time
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    ts.step(*devb[i % 4])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
