"""bench.py -- CNE+SUE training-step throughput on B200 (BASELINE.json metric: train impressions/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference's CPU path (oracle port)

One "step" = the hot path of reference trainer.py:105-120 over one batch of 64 synthetic
MIND-shaped impressions per GPU (title <= 32, abstract <= 128, history 50, K = 4, 300-d words,
gcn_layer_num 4, dropout 0.2, V = 40 000): forward, loss, backward, clip_grad_norm_(4), Adam.
N > 1: one process per GPU (torchrun), every rank its own batches (weak scaling), one NCCL
all-reduce of the flat gradient buffer per step.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = 'cne_sue_train_impressions_per_sec'
UNIT = 'impressions/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='nnr_b200', choices=['nnr_b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='impressions per GPU per step')
    ap.add_argument('--vocab', type=int, default=40000)
    ap.add_argument('--lengths', default='mind', choices=['mind', 'full', 'uniform'])
    ap.add_argument('--dropout', type=float, default=0.2)
    ap.add_argument('--cpu-sample', type=int, default=32, help='impressions per step of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true', help='skip the per-op CUDA-event breakdown')
    ap.add_argument('--gemm-detail', action='store_true', help='print the per-shape GEMM table to stderr')
    return ap.parse_args()


def workload_name(a):
    return 'cne_sue_train_step_b%d_k4_h50_t32_a128_e300_v%d_gcn4_%s_lengths' % (a.batch, a.vocab, a.lengths)


def make_config(a):
    from types import SimpleNamespace
    return SimpleNamespace(word_embedding_dim=300, vocabulary_size=a.vocab, word_threshold=3, tokenizer='MIND',
                           max_title_length=32, max_abstract_length=128, dataset='synthetic', category_num=18,
                           subCategory_num=285, category_embedding_dim=50, subCategory_embedding_dim=50,
                           dropout_rate=a.dropout, hidden_dim=200, attention_dim=200, gcn_layer_num=4,
                           no_gcn_residual=False, gcn_layer_norm=False, max_history_num=50, news_encoder='CNE',
                           user_encoder='SUE', click_predictor='dot_product', negative_sample_num=4)


# -------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line"), NVML in a thread
# -------------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
                 'hw_power_brake': 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# -------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference, bounded sample, all host threads
# -------------------------------------------------------------------------------------------------
def cpu_train_step_rate(a, sample, steps=1, warmup=0):
    from oracle import nnr_oracle as O
    from nnr_b200.synthetic import SyntheticMIND
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = O.make_config(vocabulary_size=a.vocab, dropout_rate=0.0)
    syn = SyntheticMIND(news_num=20000, vocabulary_size=a.vocab, lengths=a.lengths, seed=0)
    torch.manual_seed(0)
    p = {k: (torch.randn(s) * 0.05) for k, s in O.param_shapes(cfg).items()}
    p['news_encoder.word_embedding.weight'] = syn.word_table()
    state = {}
    times = []
    for i in range(warmup + steps):
        batch = syn.batch(sample, seed=100 + i)
        t0 = time.perf_counter()
        _, loss, grads = O.forward_backward(p, cfg, batch, lstm_impl='aten')     # ATen packed LSTM = reference path
        O.clip_and_adam(p, grads, state, i + 1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return sample / t, cores, t


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = a.cpu_sample
    rate, cores, t = cpu_train_step_rate(a, sample, steps=max(1, min(a.steps, 3)), warmup=min(a.warmup, 1))
    line = {'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(a), 'note': 'CPU only: oracle port of the reference PyTorch path'},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d impressions per step (train step: fwd+bwd+clip+Adam), packed ATen LSTM' % sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference(a)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import nnr_b200
    from nnr_b200 import ops
    from nnr_b200.synthetic import SyntheticMIND, batch_args, FIELDS
    from nnr_b200.trainer import TrainStep

    cfg = make_config(a)
    syn = SyntheticMIND(news_num=20000, vocabulary_size=a.vocab, lengths=a.lengths, seed=0)
    cfg.pretrained_word_embedding = syn.word_table()
    torch.manual_seed(1234)                                   # same init on every rank (main.py:14-15,24)
    model = nnr_b200.Model(cfg)
    model.initialize()
    model.to(dev)
    ts = TrainStep(model, lr=1e-4, gradient_clip_norm=4.0, world_size=world)
    torch.manual_seed(1000 + rank)                            # dropout streams differ per rank

    nb = 4                                                    # distinct batches per rank, rotated
    host = [syn.batch(a.batch, seed=7 * rank + i) for i in range(nb)]
    for b in host:
        for k in FIELDS:
            if torch.is_tensor(b[k]):
                b[k] = b[k].contiguous().pin_memory()
    devb = [batch_args(b, dev) for b in host]
    tok = sum(int(b['user_title_mask'].sum() + b['user_content_mask'].sum() + b['news_title_mask'].sum()
                  + b['news_content_mask'].sum()) for b in host) / nb
    slots = a.batch * 55 * 160
    h2d = sum(v.numel() * v.element_size() for k, v in host[0].items() if k in FIELDS and torch.is_tensor(v))

    def fresh(args):   # masks are mutated in place by the model (like the reference): idempotent, reuse is safe
        return args

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    for i in range(a.warmup):
        ts.step(*fresh(devb[i % nb]))
    barrier()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local_rank) as clk:
        e0.record()
        for i in range(a.steps):
            loss = ts.step(*fresh(devb[i % nb]))
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = a.batch * world * a.steps / (ms / 1e3)

    # ---- end to end: pinned host batch -> device every step, loss read back every step ----------
    for i in range(2):                                        # untimed warm-up of this path (allocator, copy engine)
        ts.step(*batch_args(host[i % nb], dev)).item()
    barrier()
    from nnr_b200.trainer import LossLog

    def run_e2e(blocking):
        """K steps from pinned host batches; every step's loss is read back to the host inside the timed region:
        blocking = float(loss) before the next step is enqueued (reference trainer.py:115); otherwise through
        trainer.LossLog (pinned slot + event, consumed while the next step runs; all K values read before the end)"""
        log = LossLog()
        vals = []
        e0.record()
        args = batch_args(host[0], dev)                       # inside the timed region: every step's batch is copied H2D
        for i in range(a.steps):
            loss = ts.step(*args)
            if not blocking:
                vals += log.push(loss, weight=a.batch)
            if i + 1 < a.steps:                               # input prefetch, like a pinned-memory DataLoader: the next
                args = batch_args(host[(i + 1) % nb], dev)    # batch's copies are enqueued before the host waits for a loss
            if blocking:
                vals.append(loss.item())
        vals += log.drain()
        e1.record()
        barrier()
        assert len(vals) == a.steps
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return a.batch * world * a.steps / (t.item() / 1e3), vals[-1]

    e2e_blocking, _ = run_e2e(True)
    e2e, loss_host = run_e2e(False)

    # ---- end to end, index-only batches: the tokenised corpus lives in HBM; per step only the
    # history / candidate ids cross PCIe, the 21 model inputs are gathered and the graph is built on the device ------
    from nnr_b200.corpus import DeviceCorpus
    corpus = DeviceCorpus.from_synthetic(syn, dev)
    ids_host = [tuple(h[k].contiguous().pin_memory() for k in ('history_ids', 'history_len', 'candidate_ids')) for h in host]
    ids_bytes = sum(v.numel() * v.element_size() for v in ids_host[0])
    for i in range(2):
        ts.step_ids(corpus, *ids_host[i % nb])
    barrier()
    e0.record()
    log = LossLog()
    n_read = 0
    for i in range(a.steps):
        loss = ts.step_ids(corpus, *ids_host[i % nb])
        n_read += len(log.push(loss, weight=a.batch))
    n_read += len(log.drain())
    assert n_read == a.steps
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ids = a.batch * world * a.steps / (t.item() / 1e3)

    # ---- per-op breakdown with CUDA events (dominant kernel -> roofline) ---------------------------
    roofline = None
    roofline_table = None
    breakdown = None
    if not a.no_profile:
        # every rank runs the two extra steps (they contain the gradient all-reduce); only rank 0 records events
        from nnr_b200 import profiler
        if rank == 0:
            with profiler.capture() as prof:
                for i in range(2):
                    ts.step(*fresh(devb[i % nb]))
                torch.cuda.synchronize()
            breakdown = prof.summary(steps=2)
            if a.gemm_detail:
                for k, v in prof.detail.items():
                    print('%-44s %8.3f ms  %4.1f calls  %7.1f TFLOP/s' % (k, v['ms'], v['calls'], v['tflops']), file=sys.stderr)
            roofline = profiler.roofline(breakdown, tokens_per_step=tok, batch=a.batch, root=ROOT)
            roofline_table = profiler.roofline_table(breakdown, ROOT)
        else:
            for i in range(2):
                ts.step(*fresh(devb[i % nb]))
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        rate, cores, tcpu = cpu_train_step_rate(a, a.cpu_sample, steps=2, warmup=1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '%d impressions per step, 1 warm-up + 2 timed train steps (fwd+bwd+clip+Adam) of the oracle port, %.1f s per step' % (a.cpu_sample, tcpu)}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'global_batch': a.batch * world, 'parallelism': 'dp%d' % world,
                       'valid_token_fraction': tok / slots, 'tokens_per_step_per_gpu': tok,
                       'l2_policy': 'per-step working set (activations+stashes, GBs) >> 126 MB L2; %d rotating batches' % nb,
                       'loss': loss_host},
            'clocks': clk.summary(),
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'loss_read': 'every step, pinned non-blocking D2H + event (trainer.LossLog), consumed while the next step '
                                 'runs; all K losses are on the host before the timed region ends',
                    'value_blocking_loss_read': e2e_blocking},
            'e2e_index_only': {'value': e2e_ids, 'unit': UNIT, 'h2d_bytes_per_step': ids_bytes, 'd2h_bytes_per_step': 4,
                               'corpus_bytes_resident': corpus.nbytes(),
                               'note': 'nnr_b200.corpus.DeviceCorpus + TrainStep.step_ids: ids in, batch gathered and graph built on the device'},
            'gpu_launches': int(launches),
            'roofline': roofline, 'roofline_by_kernel': roofline_table, 'cpu_baseline': cpu, 'breakdown_ms_per_step': breakdown}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
