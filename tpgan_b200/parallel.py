"""Data parallelism for the TP-GAN step: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), batch
sharded by rank, full replicas of G and D.  The only exchange step of the path is the gradient all-reduce (sum; the 1/world
average is folded into the optimizer's grad_scale).  The reference has no distributed code at all; this is
the green-field part of the scope table (8e).

The generator's flat gradient buffer is laid out in the order the backward plan finalises gradients, cut into contiguous
buckets; as soon as the last weight-gradient kernel of a bucket has been issued, the bucket's packed gradients are
exported into the flat buffer and its all-reduce is launched on a side stream, overlapping the rest of backward.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch


def plan_buckets(sizes: Sequence[int], ready: Sequence[int], bucket_elems: int) -> List[Tuple[int, int, int, int]]:
    """sizes[i] = elements of item i (items already sorted by `ready`, the backward index after which item i is final).
    Returns buckets (first_item, last_item_exclusive, elem_offset, elem_count) of at least bucket_elems elements each
    (the last one may be smaller); a bucket is ready at ready[last_item - 1]."""
    assert len(sizes) == len(ready) and all(ready[i] <= ready[i + 1] for i in range(len(ready) - 1))
    out, start, off, acc = [], 0, 0, 0
    for i, s in enumerate(sizes):
        acc += s
        if acc >= bucket_elems:
            out.append((start, i + 1, off, acc))
            start, off, acc = i + 1, off + acc, 0
    if acc > 0 or start < len(sizes):
        out.append((start, len(sizes), off, acc))
    return [b for b in out if b[1] > b[0]]


def allreduce_sum(t: torch.Tensor, group=None):
    """Sum all-reduce enqueued on the CURRENT stream (no host blocking: with NCCL, async_op=False only orders the stream
    after the collective).  Stream-ordered and event-free on the host side, so it can be captured into a CUDA graph."""
    import torch.distributed as dist
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=False)


class BucketReducer:
    """Bucketed, backward-overlapped gradient all-reduce over the generator's flat gradient buffer."""

    def __init__(self, trainer, bucket_mb: float = 128.0, group=None):
        from .ops import round_up
        self.tr, self.group = trainer, group
        plan, flat = trainer.plan, trainer.flat_g
        marks = plan.bwd_marks
        layers = sorted([L for L in plan.layers if L.name in marks], key=lambda L: marks[L.name])
        pname = {id(p): n for n, p in trainer.G.named_parameters()}
        sizes, ready, offs = [], [], []
        for L in layers:
            n = 0
            first = None
            for p in (L.weight, L.bias):
                if p is not None:
                    name = pname[id(p)]
                    if first is None:
                        first = flat.offsets[name]
                    n += round_up(p.numel(), 4)
            sizes.append(n)
            ready.append(marks[L.name])
            offs.append(first)
        # the flat layout follows the same order, so layer i starts at sum(sizes[:i])
        acc = 0
        for o, s in zip(offs, sizes):
            assert o == acc, "flat parameter layout must follow gradient-ready order"
            acc += s
        self.layers = layers
        self.buckets = plan_buckets(sizes, ready, int(bucket_mb * 1024 * 1024 / 4))
        from .train_step import LayerSet
        # per bucket: the deferred bias gradients of its layers + the multi-tensor gradient export
        self.sets = [LayerSet(layers[i0:i1], trainer.device, plan.bias_jobs) for (i0, i1, _, _) in self.buckets]
        self.ready = ready
        self.tail = (acc, flat.total - acc)  # parameters outside the traced layers (none for G): reduced with the last bucket
        import os
        # a high-priority stream: NCCL's CTAs are placed as soon as an SM frees up instead of behind the queued persistent grids
        self.comm = torch.cuda.Stream(priority=int(os.environ.get("TPGAN_COMM_PRIORITY", "0")))
        self.works: list = []

    def hooks(self) -> Dict[int, Callable[[], None]]:
        hk: Dict[int, Callable[[], None]] = {}
        flat = self.tr.flat_g
        for bi, (i0, i1, off, cnt) in enumerate(self.buckets):
            last = bi == len(self.buckets) - 1
            if last:
                cnt = flat.total - off

            def fn(bi=bi, off=off, cnt=cnt):
                self.sets[bi].export()
                ev = torch.cuda.Event()
                ev.record()
                self.comm.wait_event(ev)
                with torch.cuda.stream(self.comm):
                    allreduce_sum(flat.grad[off:off + cnt], self.group)
            idx = self.ready[i1 - 1]
            prev = hk.get(idx)
            hk[idx] = fn if prev is None else (lambda a=prev, b=fn: (a(), b()))
        return hk

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.comm)     # join: every bucket's all-reduce precedes the optimizer
