"""A G/D training loop in the reference's style on the B200 path.  The reference ships none: its pieces are
the modules of D_and_G_model.py, the hyper-parameters of config.py:50-85 and the helpers of UtilityMethods.py
(getOptimizer, set_requires_grad, save_model, save_optimizer) - this loop strings the drop-in versions of exactly those
together around the fused step:

    G, D            = Generator(**config.G...), Discriminator(config.D['use_batchnorm'])
    step            = TPGANTrainer.step: G forward -> WGAN-GP critic update (Adam, lr = config.train['learning_rate'])
                      -> generator update with the config.loss weights (D frozen, as set_requires_grad(D.parameters(), False)
                      would make it: no D weight gradients are computed in the G phase)
    every epoch     : save_model(G / D), save_optimizer(...)   (same file names and formats as UtilityMethods.py:58-103)

Data: any iterable of TrainDataset-style dicts (DataAndDataset.py:200-227); `SyntheticFaces` below yields raw uint8 images
+ 5-point landmarks, which the trainer normalises, crops and pyramids on the device (input_format='uint8').

    python -m tpgan_b200.train --epochs 1 --steps-per-epoch 20 --batch 32 --save-dir /tmp/tpgan
    torchrun --nproc-per-node 8 -m tpgan_b200.train --batch 32          # global batch 256, BASELINE config 4
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, Iterator

import torch

from . import config
from .D_and_G_model import Discriminator, Generator
from .train_step import TPGANTrainer
from .UtilityMethods import save_model, save_optimizer

MEAN_LANDMARKS = torch.tensor([[39.4799, 40.2799], [85.9613, 38.7062], [63.6415, 63.6473], [45.6705, 89.9648],
                               [83.9000, 88.6898]])      # D_and_G_model.py:120-128 (x, y)


class SyntheticFaces:
    """Seeded stand-in for TrainDataset: uint8 profile / frontal images, landmarks = canonical means + U(-3,3) px, identity
    label, latent z ~ U(-1,1) and the WGAN-GP interpolation coefficients.  Yields pinned host batches."""

    def __init__(self, batch: int, steps: int, seed: int = 0):
        self.batch, self.steps, self.seed = batch, steps, seed

    def __len__(self):
        return self.steps

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        g = torch.Generator().manual_seed(self.seed)
        B = self.batch
        for _ in range(self.steps):
            b = dict(img_u8=torch.randint(0, 256, (B, 128, 128, 3), generator=g, dtype=torch.uint8),
                     img_frontal_u8=torch.randint(0, 256, (B, 128, 128, 3), generator=g, dtype=torch.uint8),
                     landmarks=MEAN_LANDMARKS[None] + (torch.rand((B, 5, 2), generator=g) * 6 - 3),
                     z=torch.rand((B, config.G["zdim"]), generator=g) * 2 - 1,
                     label=torch.randint(0, config.G["num_classes"], (B,), generator=g),
                     gp_alpha=torch.rand(B, generator=g))
            pin = torch.cuda.is_available()
            yield {k: (v.contiguous().pin_memory() if pin else v.contiguous()) for k, v in b.items()}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--epochs", type=int, default=config.train["num_epochs"])
    ap.add_argument("--steps-per-epoch", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--save-dir", default="")
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--no-graphs", action="store_true")
    a = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("tpgan_b200.train needs a B200 (there is no CPU fallback)")
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)                       # identical replicas
    G = Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"]).to(device)
    D = Discriminator(config.D["use_batchnorm"]).to(device)
    trainer = TPGANTrainer(G, D, a.batch, device=device, use_dropout=True, world_size=world, use_graphs=not a.no_graphs,
                           input_format="uint8")
    # replicas are built from the SAME seed above; the device-side draws (dropout mask, background sub-sampling keys) must
    # differ per shard, as they do under DistributedDataParallel
    torch.cuda.manual_seed(0x5EED + rank)
    history = []
    for epoch in range(a.epochs):
        t0, seen, nxt = time.time(), 0, None
        it = iter(SyntheticFaces(a.batch, a.steps_per_epoch, seed=1000 * epoch + rank))
        cur = next(it, None)
        step = 0
        while cur is not None:
            nxt = next(it, None)
            m = trainer.step(cur, prefetch_next=nxt)      # next batch's host->device copies overlap this step
            history.append(m)
            seen += a.batch * world
            step += 1
            if step % a.log_every == 0 and rank == 0:
                print(f"epoch {epoch:2} step {step:5} | D {m['d_total']:8.4f} (gp {m['gp']:.4f}) | G {m['g_total']:8.4f} "
                      f"(pixel {m['pixel']:.4f} sym {m['symmetry']:.4f} tv {m['tv']:.4f} ce {m['ce']:.4f}) | "
                      f"{seen / (time.time() - t0):.1f} imgs/s", flush=True)
            cur = nxt
        if a.save_dir and rank == 0:
            save_model(G, os.path.join(a.save_dir, "G"), epoch)
            save_model(D, os.path.join(a.save_dir, "D"), epoch)
            save_optimizer(trainer.optimizer_g, G, os.path.join(a.save_dir, "G"), epoch)
            save_optimizer(trainer.optimizer_d, D, os.path.join(a.save_dir, "D"), epoch)
    if world > 1:
        dist.destroy_process_group()
    return history


if __name__ == "__main__":
    main()
