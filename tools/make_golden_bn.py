"""Writes tests/golden/generator_bn_golden.pt from the LIVE reference Generator built with use_batchnorm=True (the
constructor's default, D_and_G_model.py:351; the reference needs the F1-F4 shim of oracle/reference.py to construct it):
train-mode forward on a seeded batch of 2, the running statistics of three BatchNorm layers afterwards, every parameter's
gradient norm / sum and the full BatchNorm affine gradients for a fixed linear functional of the outputs, and the eval-mode
forward that follows.  Run in the build container:  python tools/make_golden_bn.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference as R  # noqa: E402
from oracle import step as ostep  # noqa: E402


def functional(outs, seed=11):
    """Fixed linear functional of (fake, logits, four local images): sum of seeded random weights times the outputs."""
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for o in (outs[0], outs[1], outs[3], outs[4], outs[5], outs[6]):
        w = torch.randn(o.shape, generator=g)
        total = total + (o * w.to(o.device)).sum() / o[0].numel()
    return total


def main():
    ns = R.load()
    torch.manual_seed(0)
    G = ns.DG.Generator(64, 347, True, False)
    b = ostep.make_batch(2, seed=3)
    G.train()
    outs = G(b["img"], b["left_eye"], b["right_eye"], b["nose"], b["mouth"], b["z"], False)
    functional(outs).backward()
    sd = G.state_dict()
    gold = dict(fake=outs[0].detach().clone(), logits=outs[1].detach().clone(),
                local=[outs[i].detach().clone() for i in (3, 4, 5, 6)],
                running={k: sd[k].clone() for k in ("global_pathway.conv0.0.1.running_mean", "global_pathway.conv0.0.1.running_var",
                                                    "global_pathway.deconv_8.1.running_var",
                                                    "local_pathway_nose.after_select2.0.1.running_mean")},
                grad_stats={k: (float(p.grad.norm()), float(p.grad.double().sum())) for k, p in G.named_parameters()},
                bn_grads={k: p.grad.clone() for k, p in G.named_parameters() if p.dim() == 1 and ".1." in k and p.numel() <= 512})
    G.eval()
    with torch.no_grad():
        outs = G(b["img"], b["left_eye"], b["right_eye"], b["nose"], b["mouth"], b["z"], False)
    gold["eval_fake"], gold["eval_logits"] = outs[0].clone(), outs[1].clone()
    path = os.path.join(ROOT, "tests", "golden", "generator_bn_golden.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes;", len(gold["grad_stats"]), "parameters,", len(gold["bn_grads"]), "BatchNorm vectors")


if __name__ == "__main__":
    main()
