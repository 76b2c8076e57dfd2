"""One-process-per-GPU training step for the IDEE hot path (replaces nn.DataParallel, train_synthetic.py:134-135,175-205).

All 535 892 parameters live in ONE flat fp32 buffer (the modules' nn.Parameters are views of it, so state_dict /
checkpoints are unchanged) and all gradients in a second flat buffer, so a step is

    zero flat grad -> forward -> losses -> backward -> ONE NCCL all-reduce (average) of the flat gradient -> ONE fused Adam

The batch is sharded over ranks; each rank evaluates the reference loss on its local shard and gradients are averaged,
i.e. the result equals the mean of per-shard reference steps.  Batch statistics inside the
loss (LFQ codebook entropy, BCE class weights, anomaly normaliser) are per-shard.  In the reference only the LFQ entropy is per
replica: the BCE class weights and the anomaly normaliser are computed on the gathered batch outside the DataParallel model
(train_synthetic.py:182-201), so an N-rank step here is the mean of N single-rank reference steps on the shards, not one
reference step on the global batch; a shard needs both classes present in its extreme mask (as the reference's full batch does).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .models.losses import train_step_loss


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous batch shard of rank `rank` (samples are independent through the whole path)."""
    n = t.shape[0]
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return t[rank * per:(rank + 1) * per]


class Trainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.003, lambda_anomaly=100.0,
                 process_group=None, distributed=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.lambda_anomaly = lambda_anomaly
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.step_count = 0
        self._flatten()
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        # (steps taken, lr) in device memory: the fused Adam reads them there, so one captured CUDA graph serves every step
        self.adam_state = torch.tensor([0.0, float(lr)], device=self.flat_params.device, dtype=torch.float32)
        self._graph, self._graph2 = None, None
        # Steps never run on the legacy default stream: autograd ties each parameter's gradient accumulator to the stream of its
        # first backward, and an accumulator tied to the default stream (which synchronises with every other stream) invalidates a
        # later CUDA-graph capture.  When called on the default stream the step hops onto this stream and back.
        self.stream = torch.cuda.Stream(self.flat_params.device) if self.flat_params.is_cuda else None
        if self.distributed and self.world > 1:
            dist.broadcast(self.flat_params, src=0, group=self.group)   # same weights on every rank

    @torch.no_grad()
    def _flatten(self):
        model = self.model
        packs = list(model.encoder.packs()) + list(model.cls.packs())
        in_pack = {id(p) for pk in packs for p in pk.params()}
        rest = [p for p in model.parameters() if id(p) not in in_pack]
        total = sum(pk.V * pk.P for pk in packs) + sum(p.numel() for p in rest)
        dev = next(model.parameters()).device
        flat = torch.empty(total, device=dev, dtype=torch.float32)
        off = 0
        for pk in packs:
            n = pk.V * pk.P
            pk.bind(flat[off:off + n].view(pk.V, pk.P))
            off += n
        for p in rest:
            n = p.numel()
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            off += n
        assert off == total == sum(p.numel() for p in model.parameters())
        self.flat_params = flat
        self.flat_grads = torch.zeros_like(flat)
        base = flat.data_ptr()
        self._grad_view = {}
        for p in model.parameters():
            o = (p.data_ptr() - base) // 4
            self._grad_view[id(p)] = self.flat_grads[o:o + p.numel()].view(p.shape)
            p.grad = self._grad_view[id(p)]
        # the kernels write each pack's gradient block straight into the flat buffer (no per-parameter accumulate launches)
        off = 0
        self._pack_params = []
        for pk in packs:
            n = pk.V * pk.P
            pk.grad_flat = self.flat_grads[off:off + n].view(pk.V, pk.P)
            self._pack_params.extend(pk.params())
            off += n

    def _check_flat(self):
        base, end = self.flat_params.data_ptr(), self.flat_params.data_ptr() + 4 * self.flat_params.numel()
        for p in self.model.parameters():
            if not (base <= p.data_ptr() < end):
                raise RuntimeError("Trainer: a parameter left the flat buffer (module.to()/zero_grad(set_to_none=True) after "
                                   "Trainer creation?); build the Trainer after moving the model")

    def forward_backward(self, x, mask_extreme, mask_extreme_loss):
        self.flat_grads.zero_()
        for p in self._pack_params:
            p.grad = None                      # autograd then adopts the gradient views the kernels wrote in place
        total, out = train_step_loss(self.model, x, mask_extreme, mask_extreme_loss, self.lambda_anomaly)
        total.backward()
        for p in self._pack_params:            # safety net: anything autograd cloned instead of adopting is copied back
            view = self._grad_view[id(p)]
            if p.grad is None:
                continue
            if p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
        return total, out

    def reduce_gradients(self):
        """ONE collective per step: sum the flat gradient over ranks, then average (SUM + scale works on NCCL and gloo)."""
        if self.distributed and self.world > 1:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
            self.flat_grads.mul_(1.0 / self.world)

    def optimizer_step(self):
        self.reduce_gradients()
        self.step_count += 1
        ops.adam_step_state(self.flat_params, self.flat_grads, self.exp_avg, self.exp_avg_sq, self.adam_state, self.betas[0],
                            self.betas[1], self.eps, self.weight_decay)

    def step(self, x, mask_extreme, mask_extreme_loss):
        """forward + losses + backward + gradient all-reduce + Adam; returns (loss[1] tensor on device, outputs)."""
        dev = self.flat_params.device
        if self.stream is not None and torch.cuda.current_stream(dev) == torch.cuda.default_stream(dev):
            self.stream.wait_stream(torch.cuda.default_stream(dev))
            with torch.cuda.stream(self.stream):
                total, out = self.forward_backward(x, mask_extreme, mask_extreme_loss)
                self.optimizer_step()
            torch.cuda.default_stream(dev).wait_stream(self.stream)
            return total.detach(), out
        total, out = self.forward_backward(x, mask_extreme, mask_extreme_loss)
        self.optimizer_step()
        return total.detach(), out

    # ---- learning-rate schedule / checkpointing (train_synthetic.py:127-131, utils_train.py:576-582) ----
    def set_lr(self, lr: float):
        """Schedulers (the reference steps a warm-up + cosine / step schedule every iteration) write the device-side lr."""
        self.lr = float(lr)
        self.adam_state[1] = self.lr

    def state_dict(self):
        """Optimiser state in the shape of torch.optim.Adam's essentials: flat first / second moments, step count, lr."""
        return {"step": self.step_count, "lr": self.lr, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr = float(sd["lr"])
        self.adam_state.copy_(torch.tensor([float(self.step_count), self.lr]))

    # ---- the whole step as ONE CUDA graph launch ----
    def capture(self, x, mask_extreme, mask_extreme_loss, warmup: int = 3):
        """Capture zero-grad + forward + losses + backward + all-reduce + Adam into a CUDA graph (static input buffers).

        The warm-up steps PyTorch needs before a capture run on copies of the training state, which is restored afterwards, so
        capturing does not advance training.  Returns self; use ``step_graph`` afterwards."""
        dev = self.flat_params.device
        self._static_in = tuple(torch.empty_like(t, device=dev) for t in (x, mask_extreme, mask_extreme_loss))
        for dst, src in zip(self._static_in, (x, mask_extreme, mask_extreme_loss)):
            dst.copy_(src)
        saved = (self.flat_params.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.adam_state.clone(), self.step_count)
        side = self.stream
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self._static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        from . import _lib
        before = _lib.Profile.launches
        multi = self.distributed and self.world > 1
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            if multi:
                # several ranks: the NCCL all-reduce stays OUTSIDE the graphs (one eager collective between two replays), so no
                # communicator work is ever captured: graph 1 = zero-grad + forward + losses + backward, graph 2 = Adam
                total, out = self.forward_backward(*self._static_in)
            else:
                total, out = self.step(*self._static_in)
        self._graph2 = None
        if multi:
            self._graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph2, stream=side):
                ops.adam_step_state(self.flat_params, self.flat_grads, self.exp_avg, self.exp_avg_sq, self.adam_state, self.betas[0],
                                    self.betas[1], self.eps, self.weight_decay)
        self.graph_launches = _lib.Profile.launches - before           # kernels of one replay (counted while capturing)
        self._static_out = (total.detach(), out)
        self.flat_params.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2]); self.adam_state.copy_(saved[3])
        self.step_count = saved[4]
        self._graph = graph
        return self

    def step_graph(self, x, mask_extreme, mask_extreme_loss):
        """One replay of the captured step on new inputs; the returned tensors are the graph's static outputs (valid until the next
        replay)."""
        if self._graph is None:
            raise RuntimeError("Trainer.step_graph: call capture() first")
        for dst, src in zip(self._static_in, (x, mask_extreme, mask_extreme_loss)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if self._graph2 is not None:
            self.reduce_gradients()
            self._graph2.replay()
        self.step_count += 1
        from . import _lib
        _lib.Profile.launches += self.graph_launches
        return self._static_out


class HostPrefetcher:
    """Double-buffered host -> device staging of a step's inputs on a side stream.

    The reference feeds the model from a DataLoader (train_synthetic.py:156-176) and copies each batch with ``.to(device)`` on
    the compute stream, so the copy sits in front of every step.  Here batch i+1 is copied from pinned host memory into one of
    two preallocated device slots while step i computes; ``take`` makes the compute stream wait for the copy, ``release`` tells
    the copy stream when the slot may be overwritten.  Every batch still crosses PCIe once per step."""

    def __init__(self, device, example_batch):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example_batch] for _ in range(2)]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]

    def stage(self, slot: int, host_batch):
        """Start copying ``host_batch`` (pinned tensors) into ``slot``; returns immediately."""
        with torch.cuda.stream(self.copy_stream):
            if self.free[slot] is not None:
                self.copy_stream.wait_event(self.free[slot])          # the step that read this slot has finished
            for dst, src in zip(self.slots[slot], host_batch):
                dst.copy_(src, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def take(self, slot: int):
        """Device tensors of ``slot``; the current stream waits for the copy."""
        torch.cuda.current_stream(self.device).wait_event(self.copied[slot])
        return self.slots[slot]

    def release(self, slot: int):
        """Call after the step that consumed ``slot`` has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev


class HostResults:
    """Double-buffered device -> host return of a step's results (logits, loss, driver mask) without draining the launch queue and
    without occupying the compute stream.

    The reference reads ``loss.item()`` right after every step (train_synthetic.py:176-186), which leaves the GPU idle while the
    host enqueues the next step.  Here step i's results are snapshotted into device staging slot i & 1 behind the step on the
    compute stream (a few microseconds; this is also where a dtype change such as int64 -> uint8 happens), copied to pinned host
    memory by a side stream (``put``), and read by the host (``get``) one step later, after step i+1 has been enqueued: every
    step's results still reach the host, the PCIe transfer overlaps the next step's kernels."""

    def __init__(self, device, example_results):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in example_results] for _ in range(2)]
        self.stage = [[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example_results] for _ in range(2)]
        self.snap = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.used = [False, False]

    def put(self, slot: int, results):
        """Snapshot ``results`` on the current stream and start their device -> host copies on the side stream; returns
        immediately.  Slot reuse is safe once ``get`` of the same slot has returned (it waits for the copies)."""
        cur = torch.cuda.current_stream(self.device)
        if self.used[slot]:
            cur.wait_event(self.done[slot])                       # the previous transfer out of this staging slot has finished
        self.used[slot] = True
        for dst, src in zip(self.stage[slot], results):
            dst.copy_(src.detach().reshape(dst.shape), non_blocking=True)
        self.snap[slot].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.snap[slot])
            for dst, src in zip(self.slots[slot], self.stage[slot]):
                dst.copy_(src, non_blocking=True)
            self.done[slot].record(self.copy_stream)

    def get(self, slot: int):
        """Pinned host tensors of ``slot`` (blocks until its copies have landed)."""
        self.done[slot].synchronize()
        return self.slots[slot]
