"""Pretrain.py's training iteration (Pretrain.py:159-181) for MobileNetV2 as ONE flat launch schedule on the device:

    stage images (NCHW -> NHWC, tf32) -> MobileNetV2 forward (training-mode BatchNorm) -> batched MultiTaskLoss (loss value,
    assignment, dL/dlocations, dL/dclassifications in one launch) -> backward -> gradient export into the flat .grad buffer
    -> [NCCL all-reduce] -> SGD-Nesterov (UtilityMethods.py:30, config.py:31-35) -> weight re-pack

with no host synchronisation inside (the reference syncs per predicted point: `.item()` at MobileNetV2.py:411-430), captured
into a CUDA graph on request.  Data parallelism: full replica per rank, batch sharded by rank, one all-reduce
of the 30 MB flat gradient; BatchNorm statistics stay rank-local, exactly as nn.BatchNorm2d under DistributedDataParallel
without SyncBN (the reference has none) - an N-GPU run equals a 1-GPU run only at equal per-GPU batch.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops
from .engine import GradArena, Plan
from .MobileNetV2 import MobileNetV2
from .train_step import Eager, FlatOptimizer, FlatParams, GraphRunner, LayerSet

# config.py:25-35 (pretrain['loss'], optimizer_param); the reference module itself is not imported by the product
LOSS = dict(alpha=30.0, beta=0.1, ratio_non_background=5.0, distance_threshold_ratio=0.1)
OPTIM = dict(learning_rate=5e-4, momentum=0.9, nesterov=True, weight_decay=5e-4)
LR_MILESTONES, LR_GAMMA = (10, 20, 30), 0.1     # config.py:16-18, Pretrain.py:117-121 (MultiStepLR)


class _FlatSGDTrainer:
    """What the two pre-training trainers share: flat parameter / gradient / momentum buffers, the traced plan with
    direct parameter gradients, the launch schedule (stage -> forward -> loss -> backward -> export -> [all-reduce] ->
    SGD-Nesterov -> re-pack), CUDA-graph replay, MultiStepLR and the torch.optim-compatible optimizer view.  Subclasses
    provide _trace_model(plan) (returns the output tensors whose gradients the loss writes), load_inputs, _stage, _loss."""

    def _setup(self, model, B: int, image_hw, device, exact: bool, world_size: int, group, use_graphs: bool):
        self.model, self.B, self.device = model, B, torch.device(device)
        self.hw = tuple(image_hw)
        self.world_size, self.group, self.use_graphs = world_size, group, use_graphs
        model.train()
        self.flat = FlatParams(model)                      # params + .grad as views of flat buffers (m = momentum buffer)
        self.lr = float(OPTIM["learning_rate"])            # live learning rate: host copy + the device word the SGD kernel reads
        self.lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=self.device)
        # what getOptimizer(model.parameters(), 'SGD')() would hold; save_optimizer(tr.optimizer, model, dir, epoch) works
        self.optimizer = FlatOptimizer(self.flat, model, "sgd", dict(lr=OPTIM["learning_rate"], momentum=OPTIM["momentum"],
                                                                     weight_decay=OPTIM["weight_decay"],
                                                                     nesterov=OPTIM["nesterov"]),
                                       lr_get=lambda: self.lr, lr_set=self._set_lr)
        self.epoch = 0
        plan = Plan(self.device, training=True, need_wgrad=True, exact=exact, defer_bias=True)
        plan.direct_grads = True
        self.plan = plan
        H, W = self.hw
        self.x = plan.new(B, H, W, 3, name="images", requires_grad=False)
        for t in self._trace_model(plan):
            plan.seed_grad(t)
        plan.trace_backward()
        self.set = LayerSet(plan.layers, self.device, plan.bias_jobs)
        self.set.repack()
        self.sums = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.inp: Optional[Dict[str, torch.Tensor]] = None
        self._sched: Dict[tuple, object] = {}
        self.steps = 0


class PretrainTrainer(_FlatSGDTrainer):
    """step(images (B,3,H,W) in [-1,1], labels (B,8) = 4 ground-truth points (x,y), u (B,n) optional sub-sampling keys)."""

    def __init__(self, model: MobileNetV2, B: int, image_hw=(128, 128), device="cuda", exact: bool = True,
                 world_size: int = 1, group=None, use_graphs: bool = False):
        """exact=True (default): dense convolutions in the fp32-accurate 3xTF32 split mode - the production mode of this
        network (see MobileNetV2's docstring: within 1e-3 of the fp32 reference); exact=False: single-pass TF32."""
        self._setup(model, B, image_hw, device, exact, world_size, group, use_graphs)
        self.n = self.loc.act.c // 2
        assert self.cls.act.c == 5 * self.n
        self.labels = torch.zeros((B, self.n), dtype=torch.int32, device=self.device)

    def _trace_model(self, plan: Plan):
        self.loc, self.cls = self.model.trace(plan, self.x)
        return [self.loc, self.cls]

    # ---- inputs live in static device buffers so that every pointer of the schedule is fixed
    def load_inputs(self, images: torch.Tensor, labels: torch.Tensor, u: Optional[torch.Tensor]):
        if self.inp is None:
            self.inp = dict(images=torch.empty((self.B, 3) + self.hw, dtype=torch.float32, device=self.device),
                            labels=torch.empty((self.B, 8), dtype=torch.float32, device=self.device),
                            u=torch.empty((self.B, self.n), dtype=torch.float32, device=self.device))
        self.inp["images"].copy_(images, non_blocking=True)
        self.inp["labels"].copy_(labels.reshape(self.B, 8), non_blocking=True)
        if u is not None:
            self.inp["u"].copy_(u, non_blocking=True)
        self._draw_u = u is None

    # ---- the NEXT batch's host->device copies on a side stream while this step runs (pinned host tensors)
    def prefetch(self, batch):
        """batch = the tuple a later step() will be called with (the same object): staged into device buffers on a copy
        stream now, picked up with device-to-device copies then."""
        if getattr(self, "_pf_stream", None) is None:
            self._pf_stream = torch.cuda.Stream(device=self.device)
            self._pf_event = torch.cuda.Event()
            self._pf_picked = None
            self._pf_bufs = [None if t is None else torch.empty_like(t, device=self.device) for t in batch]
        if self._pf_picked is not None:
            self._pf_stream.wait_event(self._pf_picked)    # staging buffers are free once the last pick-up has run
        with torch.cuda.stream(self._pf_stream):
            for dst, src in zip(self._pf_bufs, batch):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self._pf_event.record(self._pf_stream)
        self._pf_batch = batch

    def _take_prefetched(self, batch):
        """If `batch` is the object handed to prefetch(), return its staged device copies (after waiting for the copy)."""
        if getattr(self, "_pf_batch", None) is not batch or batch is None:
            return batch
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._pf_event)
        self._pf_batch = None
        return tuple(self._pf_bufs)

    def _mark_picked(self):
        if getattr(self, "_pf_stream", None) is not None:
            if self._pf_picked is None:
                self._pf_picked = torch.cuda.Event()
            self._pf_picked.record(torch.cuda.current_stream(self.device))

    def _stage(self):
        self.x.act.from_nchw(self.inp["images"], round_tf32=not self.plan.exact)
        if self._draw_u:
            self.inp["u"].uniform_()        # the torch.multinomial draw of MobileNetV2.py:505 as per-point keys (philox)

    def _loss(self):
        a = LOSS
        H, W = self.hw
        dloc, dcls = self.plan.grad_act(self.loc).buf, self.plan.grad_act(self.cls).buf
        ops.multitask_loss(self.loc.act.buf, self.cls.act.buf, self.inp["labels"], self.inp["u"], self.n,
                           self.loc.act.buf.shape[3], self.cls.act.buf.shape[3],
                           int(a["distance_threshold_ratio"] * self.n), float(W), float(H), a["alpha"], a["beta"],
                           a["ratio_non_background"], 1.0 / self.B, dloc, dcls, self.labels, self.sums)

    def _schedule(self, optimize: bool) -> List[Callable]:
        sch: List[Callable] = [self._stage]
        sch += self.plan.fwd
        sch += [lambda: self.sums.zero_(), self._loss, GradArena.get(self.device).zero, lambda: self.flat.grad.zero_()]
        sch += self.plan.bwd
        sch.append(self.set.export)
        if self.world_size > 1:
            sch.append(Eager(self._allreduce))
        if optimize:
            o = OPTIM
            sch.append(lambda: ops.sgd_step(self.flat.data, self.flat.grad, self.flat.m, self.lr_dev, o["momentum"],
                                            o["weight_decay"], o["nesterov"], 1.0 / self.world_size))
            sch.append(self.set.repack)
        return sch

    def _allreduce(self):
        import torch.distributed as dist
        dist.all_reduce(self.flat.grad, group=self.group)

    def step(self, images: torch.Tensor, labels: torch.Tensor, u: Optional[torch.Tensor] = None, optimize: bool = True,
             read_metrics: bool = True, batch=None, prefetch_next=None):
        """batch: alternatively the (images, labels[, u]) tuple previously handed to prefetch(); prefetch_next: the tuple of
        the NEXT step, whose host->device copies are started on the copy stream once this step's launches are enqueued."""
        if batch is not None:
            staged = self._take_prefetched(batch)
            images, labels = staged[0], staged[1]
            u = staged[2] if len(staged) > 2 else None
        self.load_inputs(images, labels, u)
        if batch is not None:
            self._mark_picked()
        key = (optimize, bool(getattr(self, "_draw_u", False)))   # _stage branches on _draw_u; a captured graph freezes it
        if key not in self._sched:
            sch = self._schedule(optimize)
            self._sched[key] = GraphRunner(sch) if self.use_graphs else sch
        sch = self._sched[key]
        if self.use_graphs:
            sch.run()
        else:
            for f in sch:
                f()
        self.steps += 1
        inv = getattr(getattr(self, "net", None), "invalidate_folded", None)
        if inv is not None:
            inv()       # parameters / running statistics were written through raw pointers: eval-mode folded copies are stale
        if prefetch_next is not None:
            self.prefetch(prefetch_next)
        return self.read_metrics() if read_metrics else None

    def read_metrics(self) -> Dict[str, float]:
        s = self.sums.cpu().tolist()
        return dict(loss=s[0], location=s[1], classification=s[2])

    def outputs(self):
        """(locations (B,n,2), classifications (B,n,5)) of the last step (copies)."""
        n = self.n
        return (self.loc.act.buf.view(self.B, -1)[:, :2 * n].reshape(self.B, n, 2).clone(),
                self.cls.act.buf.view(self.B, -1)[:, :5 * n].reshape(self.B, n, 5).clone())

    def end_epoch(self):
        """learning_rate_scheduler.step() (Pretrain.py:296): MultiStepLR(milestones, gamma)."""
        self.set_epoch(self.epoch + 1)

    def set_epoch(self, epoch: int):
        """Resume: restore the MultiStepLR position (the scheduler's last_epoch) and the device-resident learning rate."""
        self.epoch = int(epoch)
        k = sum(1 for m in LR_MILESTONES if self.epoch >= m)
        self._set_lr(OPTIM["learning_rate"] * (LR_GAMMA ** k))

    def _set_lr(self, lr: float):
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    def sync_buffers(self):
        """num_batches_tracked of every BatchNorm2d (state_dict parity with nn.BatchNorm2d in train mode)."""
        for m in self.model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)) and m.num_batches_tracked is not None:
                m.num_batches_tracked.fill_(self.steps)


class ClassifierTrainer(PretrainTrainer):
    """Pre-training of the feature extractor as an identity classifier (BASELINE config 5, "ResNet backbones"; what
    FeatureExtract.py:5-41 builds: a backbone + Linear(., num_of_output_classes)): ResNet18-128 forward with batch-statistics
    BatchNorm, softmax cross-entropy, backward, SGD-Nesterov - the same fused schedule as PretrainTrainer.
    step(images (B,3,128,128) in [-1,1], labels (B,) int64)."""

    def __init__(self, model, B: int, image_hw=(128, 128), device="cuda", exact: bool = False, world_size: int = 1,
                 group=None, use_graphs: bool = False):
        net = getattr(model, "base_model", model)      # FeatureExtractModel or the ResNet18 itself
        self.net = net
        self._setup(model, B, image_hw, device, exact, world_size, group, use_graphs)
        self.num_classes = self.logits.act.c

    def _trace_model(self, plan: Plan):
        self.logits, self.fc0, self.pooled = self.net.trace_train(plan, self.x)
        return [self.logits]

    def load_inputs(self, images: torch.Tensor, labels: torch.Tensor, u=None):
        if self.inp is None:
            self.inp = dict(images=torch.empty((self.B, 3) + self.hw, dtype=torch.float32, device=self.device),
                            labels=torch.empty((self.B,), dtype=torch.int64, device=self.device))
        self.inp["images"].copy_(images, non_blocking=True)
        self.inp["labels"].copy_(labels.reshape(self.B), non_blocking=True)

    def _stage(self):
        self.x.act.from_nchw(self.inp["images"], round_tf32=not self.plan.exact)

    def _loss(self):
        ops.softmax_ce(self.logits.act, self.inp["labels"], self.plan.grad_act(self.logits), 1.0 / self.B, self.sums[0:1])

    def read_metrics(self) -> Dict[str, float]:
        return dict(loss=float(self.sums[0].cpu()) / self.B)

    def outputs(self):
        return self.logits.act.buf.view(self.B, -1)[:, :self.num_classes].clone()
