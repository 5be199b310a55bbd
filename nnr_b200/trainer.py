"""The training-step contract of reference trainer.py:64-66,105-120 (single GPU) and :285-300 (DDP).

    logits = model(*21 tensors); loss = mean(-log_softmax(logits)[:, 0]); zero_grad; backward;
    clip_grad_norm_(params, 4); Adam(lr=1e-4).step()

B200-native layout: all parameters live in ONE flat fp32 buffer (``param.data`` are views), gradients
in a second flat buffer (``param.grad`` are views), Adam moments in two more.  Data parallelism is
one process per GPU and exactly one ``ncclAllReduce`` over the flat gradient buffer per step
(the reference's DDP issues bucketed all-reduces, trainer.py:219,297); the 1/world averaging is
folded into the fused clip+Adam kernel (``nnr_flat_clip_adam``).
"""
import torch
import torch.distributed as dist

from . import engine, ops


def negative_log_softmax(logits):
    """trainer.py:64-66"""
    return (-torch.log_softmax(logits, dim=1).select(dim=1, index=0)).mean()


class TrainStep:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, gradient_clip_norm=4.0, process_group=None,
                 world_size=None):
        self.model = model
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, gradient_clip_norm
        self.pg = process_group
        if world_size is None:
            world_size = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world_size = world_size
        params = [p for p in model.parameters() if p.requires_grad]
        dev = params[0].device          # flat buffers / all-reduce are device agnostic; the optimizer kernel is CUDA only
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]            # keep every slice 16-byte aligned
        total = sum(sizes)
        self.flat = torch.zeros(total, device=dev)
        self.gflat = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        o = 0
        for p, s in zip(params, sizes):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
            p.grad = self.gflat[o:o + n].view(p.shape)
            p._nnr_flat_grad = True          # engine._param_grads adds straight into these views
            o += s
        self.params = params
        # operand planes of every GEMM weight matrix, refreshed by one launch after each optimizer step (instead of one
        # small split launch per weight per step); embedding tables are gathered, not multiplied
        self._gemm_weights = [p for n, p in model.named_parameters()
                              if p.requires_grad and p.dim() == 2 and p.is_cuda and 'embedding' not in n and '_lstm.' not in n]
        self._gemm_weights = list({id(p): p for p in self._gemm_weights}.values())
        self._split = ops.SplitMany(self._gemm_weights) if dev.type == 'cuda' else None
        self._refresh_weight_planes()
        self.step_count = 0
        self.grad_norm = torch.zeros(1, device=dev)

    def _refresh_weight_planes(self):
        engine.weights_changed()
        if self._split is not None and self._split.n:
            self._split.refresh()
            for w, pl in zip(self._gemm_weights, self._split.planes):
                engine.install_weight_planes(w, pl)

    def zero_grad(self):
        self.gflat.zero_()
        for p in self.params:                                          # keep .grad bound to the flat views
            if p.grad is None or p.grad.data_ptr() < self.gflat.data_ptr() or \
                    p.grad.data_ptr() >= self.gflat.data_ptr() + self.gflat.numel() * 4:
                raise RuntimeError('parameter .grad was rebound; use TrainStep.zero_grad() only')

    def step(self, *batch):
        """one full training step; returns the (device) loss tensor without synchronising"""
        if not self.model.training:
            self.model.train()
        logits = self.model(*batch)
        loss = negative_log_softmax(logits)
        self.gflat.zero_()
        loss.backward()
        self.optimizer_step()
        return loss

    def step_ids(self, corpus, history_ids, history_len, candidate_ids):
        """index-only training step: the batch is gathered from a ``corpus.DeviceCorpus`` on the device"""
        return self.step(*corpus.batch_from_ids(history_ids, history_len, candidate_ids))

    def reduce_gradients(self):
        """the single collective of the step: SUM over ranks of the flat gradient buffer (the 1/world average
        of DDP, trainer.py:219, is applied inside the fused clip+Adam kernel as grad_scale)"""
        if self.world_size > 1:
            dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self):
        self.reduce_gradients()
        if self.flat.device.type != 'cuda':
            raise RuntimeError('nnr_b200.TrainStep.optimizer_step needs CUDA (nnr_flat_clip_adam has no CPU path)')
        self.step_count += 1
        ops.flat_clip_adam(self.flat, self.gflat, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                           self.eps, self.max_norm, 1.0 / self.world_size, self.step_count, self.grad_norm)
        self._refresh_weight_planes()   # the kernel updates the parameters through raw pointers: re-split them (one launch)


class LossLog:
    """Per-step loss read-back without stalling the launch queue.

    The reference adds ``float(loss) * batch`` to the epoch loss every step (trainer.py:115,295), which blocks the host
    until the step has finished before the next one is enqueued.  Here every step's loss is copied into a pinned host
    slot with a non-blocking D2H copy + an event; ``push`` returns the values whose copies have completed (normally the
    previous step's), ``drain`` waits for the rest.  Every step is still read exactly once and the epoch sum is the
    same -- the host just consumes step i while the device runs step i+1."""

    def __init__(self, slots=8):
        self.buf = torch.empty(slots, dtype=torch.float32).pin_memory()
        self.pending = []          # (slot, event, weight) in issue order
        self.slots = slots
        self.next = 0
        self.total = 0.0           # sum of weight * loss over the consumed steps (the reference's epoch_loss)
        self.count = 0

    def _consume(self, block):
        out = []
        while self.pending and (block or self.pending[0][1].query()):
            slot, ev, wt = self.pending.pop(0)
            ev.synchronize()
            v = float(self.buf[slot])
            self.total += v * wt
            self.count += 1
            out.append(v)
        return out

    def push(self, loss, weight=1.0, lag=1):
        """enqueue the read of this step's (device) loss; returns the losses that are complete, waiting only if more than
        ``lag`` reads are outstanding"""
        if len(self.pending) >= self.slots:
            self._consume_one()
        slot = self.next
        self.next = (self.next + 1) % self.slots
        self.buf[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((slot, ev, weight))
        out = []
        while len(self.pending) > lag:
            out += self._consume_one()
        return out + self._consume(False)

    def _consume_one(self):
        slot, ev, wt = self.pending.pop(0)
        ev.synchronize()
        v = float(self.buf[slot])
        self.total += v * wt
        self.count += 1
        return [v]

    def drain(self):
        return self._consume(True)


def shard_batch(batch, rank, world_size):
    """Reference DDP sharding (trainer.py:218,256-258): rank r takes impressions [r*B/W, (r+1)*B/W) of a global
    batch whose size is divisible by the world size (config.py:116).  Works on dicts or sequences of tensors;
    non-tensor entries (the unused entity fields may be None) are passed through."""
    def cut(v):
        if not torch.is_tensor(v):
            return v
        B = v.shape[0]
        if B % world_size:
            raise ValueError('batch size %d not divisible by world size %d' % (B, world_size))
        per = B // world_size
        return v[rank * per:(rank + 1) * per]
    if isinstance(batch, dict):
        return {k: cut(v) for k, v in batch.items()}
    return [cut(v) for v in batch]
