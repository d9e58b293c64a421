"""The reference's Pretrain.py loop (Pretrain.py:76-309) on the B200 path: MobileNetV2-SSD landmark pre-training with
MultiTaskLoss, SGD-Nesterov (getOptimizer 'SGD', config.py:14,31-35), MultiStepLR (config.py:16-18), the decoder + accuracy
metric every step (Pretrain.py:165-168), a validation pass in eval mode every `log_step_of_batchs` steps (:184-236) and
save_model / save_optimizer at the end of every epoch (:301-302).

What differs, deliberately: batches instead of the reference's batch_size 1 (config.py:12 - its loss cannot take more),
the whole iteration is one launch schedule (PretrainTrainer), nothing is synchronised per step except the metric read,
and the data source is any iterable of (images (B,3,H,W) in [-1,1], labels (B,8)) - `SyntheticLandmarks` below when no
dataset is at hand (PretrainDataset needs the CelebA files of config.py:4-5; file IO is outside this repository's scope).

    python -m tpgan_b200.Pretrain --epochs 1 --steps-per-epoch 50 --batch 32 --log-dir /tmp/pretrain
    torchrun --nproc-per-node 8 -m tpgan_b200.Pretrain --batch 32          # global batch 256, BASELINE config 5
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Iterable, Iterator, Tuple

import torch

from .MobileNetV2 import MobileNetV2, MultiTaskDecoder, MultiTaskLoss
from .pretrain_step import PretrainTrainer
from .UtilityMethods import save_model, save_optimizer

# config.py:1-27 (pretrain[...]); the reference's config module is not imported by the product
pretrain = {"model_name": "MobileNetV2", "optimizer": "SGD", "use_learning_rate_scheduler": True,
            "learning_rate_scheduler_milestone": [10, 20, 30], "learning_rate_scheduler_gamma": 0.1, "num_epochs": 5,
            "log_step_of_batchs": 200}


class SyntheticLandmarks:
    """Seeded stand-in for PretrainDataset (DataAndDataset.py:58-177): faces ~ U(-1,1), four ground-truth points = the
    canonical landmark means (D_and_G_model.py:120-128) + U(-3,3) px.  Yields pinned host batches."""

    MEANS = torch.tensor([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [64.7803, 89.3250]])

    def __init__(self, batch: int, steps: int, seed: int = 0, hw: Tuple[int, int] = (128, 128)):
        self.batch, self.steps, self.seed, self.hw = batch, steps, seed, hw

    def __len__(self):
        return self.steps

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.steps):
            images = torch.rand((self.batch, 3) + self.hw, generator=g) * 2 - 1
            labels = (self.MEANS[None] + (torch.rand((self.batch, 4, 2), generator=g) * 6 - 3)).reshape(self.batch, 8)
            if torch.cuda.is_available():
                images, labels = images.pin_memory(), labels.pin_memory()
            yield images, labels


def validate(model: MobileNetV2, loss_fn: MultiTaskLoss, decoder: MultiTaskDecoder, loader: Iterable, device) -> Tuple[float, float]:
    """Pretrain.py:184-236: eval mode (running BatchNorm statistics), no gradients; mean loss and accuracy."""
    model.eval()
    losses, accs = [], []
    with torch.no_grad():
        for images, labels in loader:
            images, labels = images.to(device, non_blocking=True), labels.to(device, non_blocking=True)
            loc, cls = model(images, use_dropout=False)
            losses.append(loss_fn(loc, cls, labels, (images.size(2), images.size(3))))
            accs.append(decoder.decode(loc, cls, labels)[3].mean())
    model.train()
    return float(torch.stack(losses).mean()), float(torch.stack(accs).mean())


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--epochs", type=int, default=pretrain["num_epochs"])
    ap.add_argument("--steps-per-epoch", type=int, default=100)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--log-dir", default="")
    ap.add_argument("--log-every", type=int, default=pretrain["log_step_of_batchs"])
    ap.add_argument("--val-steps", type=int, default=2)
    ap.add_argument("--no-graphs", action="store_true")
    a = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("tpgan_b200.Pretrain needs a B200 (there is no CPU fallback)")
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)                       # identical replicas
    model = MobileNetV2().to(device)
    trainer = PretrainTrainer(model, a.batch, device=device, world_size=world, use_graphs=not a.no_graphs)
    # replicas are built from the SAME seed above; the device-side draws (dropout mask, background sub-sampling keys) must
    # differ per shard, as they do under DistributedDataParallel
    torch.cuda.manual_seed(0x5EED + rank)
    loss_fn, decoder = MultiTaskLoss(), MultiTaskDecoder()
    history = []
    for epoch in range(a.epochs):
        t0, seen = time.time(), 0
        it = iter(SyntheticLandmarks(a.batch, a.steps_per_epoch, seed=1000 * epoch + rank))
        cur, step = next(it, None), 0
        if cur is not None:
            trainer.prefetch(cur)
        while cur is not None:
            nxt = next(it, None)
            # forward, loss, backward, SGD: one schedule; the next batch's host->device copies overlap it
            m = trainer.step(None, None, batch=cur, prefetch_next=nxt)
            acc = float(decoder.decode(*trainer.outputs(), trainer.inp["labels"])[3].mean())   # Pretrain.py:165-168
            seen += a.batch * world
            history.append((m["loss"], acc))
            step += 1
            if step % a.log_every == 0 and rank == 0:
                vl, va = validate(model, loss_fn, decoder, SyntheticLandmarks(a.batch, a.val_steps, seed=777), device)
                print(f"===== epoch: {epoch:2}, step: {step:6} / {a.steps_per_epoch} =====\n train_loss: {m['loss']:6.4f}, "
                      f"train_accuracy: {acc:.4f}\nval_loss: {vl:6.4f}, val_accuracy {va:.4f}\n"
                      f"{seen / (time.time() - t0):.1f} imgs/s", flush=True)
            cur = nxt
        if pretrain["use_learning_rate_scheduler"]:
            trainer.end_epoch()                                               # learning_rate_scheduler.step(), Pretrain.py:296
        if a.log_dir and rank == 0:
            trainer.sync_buffers()
            save_model(model, a.log_dir, epoch)                               # Pretrain.py:301-302
            save_optimizer(trainer.optimizer, model, a.log_dir, epoch)
    if world > 1:
        dist.destroy_process_group()
    return history


if __name__ == "__main__":
    main()
