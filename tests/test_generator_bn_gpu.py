"""Generator(zdim, num_classes, use_batchnorm=True, ...) - the reference constructor's DEFAULT (D_and_G_model.py:351; config.py
turns it off) - against golden vectors recorded from the live reference (tools/make_golden_bn.py; the reference needs the
F1-F4 shim to construct this variant): conv / deconv (no bias) -> batch-statistics BatchNorm2d -> (Leaky)ReLU in every
factory stack of the two pathways, 55 BatchNorm layers, outputs written straight into the concat buffers.

fp32-exact mode pins it: outputs <= 1e-3 (training-mode BatchNorm over a batch of 2 amplifies round-off: 2 x 5x5 = 50 samples
per channel in the local pathways' bottleneck), running statistics <= 1e-3, gradients: every BatchNorm affine gradient and
the per-parameter norms <= 5e-2 overall.  TF32: outputs sanity-bounded (0.15)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "generator_bn_golden.pt")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _functional(outs, seed=11):      # tools/make_golden_bn.py::functional
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for o in (outs[0], outs[1], outs[3], outs[4], outs[5], outs[6]):
        w = torch.randn(o.shape, generator=g)
        total = total + (o * w.to(o.device)).sum() / o[0].numel()
    return total


@pytest.mark.parametrize("exact", [True, False])
def test_generator_with_batchnorm_vs_live_reference_goldens(exact):
    import tpgan_b200.D_and_G_model as M
    from oracle import step as ostep
    gold = torch.load(GOLD, weights_only=False)
    M.EXACT_MODE = exact
    try:
        torch.manual_seed(0)
        G = M.Generator(64, 347, True, False)          # seeded init is bit-identical to the reference's (548 state_dict keys)
        assert len(G.state_dict()) == 548
        G.cuda().train()
        b = {k: v.cuda() for k, v in ostep.make_batch(2, seed=3).items()}
        outs = G(b["img"], b["left_eye"], b["right_eye"], b["nose"], b["mouth"], b["z"], False)
        tol = 1e-3 if exact else 0.15
        assert rel(outs[0], gold["fake"]) < tol and rel(outs[1], gold["logits"]) < tol, (rel(outs[0], gold["fake"]),
                                                                                         rel(outs[1], gold["logits"]))
        for o, w in zip((outs[3], outs[4], outs[5], outs[6]), gold["local"]):
            assert rel(o, w) < tol
        if not exact:
            return
        sd = G.state_dict()
        for k, v in gold["running"].items():
            assert rel(sd[k], v) < 1e-3, k
        assert int(sd["global_pathway.conv0.0.1.num_batches_tracked"]) == 1
        _functional(outs).backward()
        num = den = 0.0
        for k, p in G.named_parameters():
            n_ref, s_ref = gold["grad_stats"][k]
            assert p.grad is not None, k
            num += (float(p.grad.norm()) - n_ref) ** 2
            den += n_ref ** 2
        assert math.sqrt(num / den) < 5e-2, math.sqrt(num / den)
        pg = dict(G.named_parameters())
        num = den = 0.0
        for k, g in gold["bn_grads"].items():
            num += float((pg[k].grad.double().cpu() - g.double()).pow(2).sum())
            den += float(g.double().pow(2).sum())
        assert math.sqrt(num / den) < 5e-2, math.sqrt(num / den)
        G.eval()
        with torch.no_grad():
            oe = G(b["img"], b["left_eye"], b["right_eye"], b["nose"], b["mouth"], b["z"], False)
        assert rel(oe[0], gold["eval_fake"]) < 2e-3 and rel(oe[1], gold["eval_logits"]) < 2e-3
    finally:
        M.EXACT_MODE = None


def test_fused_trainer_rejects_batchnorm_critic():
    import tpgan_b200.D_and_G_model as M
    from tpgan_b200.train_step import TPGANTrainer
    G, D = M.Generator(64, 347, False, False).cuda(), M.Discriminator(True).cuda()
    with pytest.raises(NotImplementedError, match="module API"):
        TPGANTrainer(G, D, 1)


STEP_GOLD = os.path.join(os.path.dirname(__file__), "golden", "step_bn_golden.pt")


@pytest.mark.parametrize("exact", [True, False])
def test_fused_step_with_batchnorm_generator_vs_live_reference_goldens(exact):
    """TPGANTrainer over Generator(use_batchnorm=True) + the config's BatchNorm-free Discriminator: one fused step at fixed
    weights against the oracle step run on the LIVE reference modules (tools/make_golden_bn_step.py): the 11 loss scalars,
    every BatchNorm affine gradient, per-parameter gradient norms of both networks, the running statistics after the step's
    single generator forward.  Exact (3xTF32) mode pins it; tf32 is sanity-bounded (training-mode BatchNorm over a batch of
    2 amplifies the operand rounding, see the module docstring)."""
    import tpgan_b200.D_and_G_model as M
    from oracle import step as ostep
    from tpgan_b200 import _lib
    from tpgan_b200.train_step import TPGANTrainer
    gold = torch.load(STEP_GOLD, weights_only=False)
    torch.manual_seed(0)
    G = M.Generator(64, 347, True, False)
    D = M.Discriminator(False)
    G.cuda()
    D.cuda()
    B = 2
    b = ostep.make_batch(B, seed=3)
    tr = TPGANTrainer(G, D, B, exact=exact)
    m = tr.step({k: v.cuda() for k, v in b.items()}, optimize=False)
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    tol_m = 2e-3 if exact else 0.2
    for k, v in gold["metrics"].items():
        assert abs(m[k] - v) <= tol_m * abs(v) + 1e-4, (k, m[k], v)
    if not exact:
        return
    tr.sync_buffers()
    sd = G.state_dict()
    for k, v in gold["running"].items():
        assert rel(sd[k], v) < 1e-3, k
    assert int(sd["global_pathway.conv0.0.1.num_batches_tracked"]) == gold["num_batches_tracked"] == 1
    for net, stats in ((G, gold["g_grad_stats"]), (D, gold["d_grad_stats"])):
        num = den = 0.0
        for k, p in net.named_parameters():
            n_ref, _ = stats[k]
            assert p.grad is not None, k
            num += (float(p.grad.norm()) - n_ref) ** 2
            den += n_ref ** 2
        assert math.sqrt(num / den) < 5e-2, math.sqrt(num / den)
    pg = dict(G.named_parameters())
    num = den = 0.0
    for k, g in gold["bn_grads"].items():
        num += float((pg[k].grad.double().cpu() - g.double()).pow(2).sum())
        den += float(g.double().pow(2).sum())
    assert math.sqrt(num / den) < 5e-2, math.sqrt(num / den)


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_fused_step_with_batchnorm_generator_trains_and_replays_as_graph(dtype):
    """Three optimizer steps of the BatchNorm generator through CUDA graphs (tf32 and bf16 operands - the BatchNorm kernels
    read and write the fp32 copies, the engine re-casts their outputs for the tensor-core consumers): the first step's losses
    within the format's sanity bound of the live-reference golden, finite losses, parameters and running statistics move, the
    affine parameters of a BatchNorm layer receive Adam updates, num_batches_tracked follows the step count."""
    import tpgan_b200.D_and_G_model as M
    from oracle import step as ostep
    from tpgan_b200.train_step import TPGANTrainer
    gold = torch.load(STEP_GOLD, weights_only=False)
    torch.manual_seed(0)
    G, D = M.Generator(64, 347, True, False).cuda(), M.Discriminator(False).cuda()
    B = 2
    tr = TPGANTrainer(G, D, B, use_graphs=True, dtype=dtype)
    m0 = tr.step({k: v.cuda() for k, v in ostep.make_batch(B, seed=3).items()})
    for k in ("pixel", "local", "symmetry", "tv", "ce", "g_total"):     # the terms that do not hinge on the critic's sign
        assert abs(m0[k] - gold["metrics"][k]) <= 0.25 * abs(gold["metrics"][k]) + 1e-3, (k, m0[k], gold["metrics"][k])
    bn = G.global_pathway.conv0[0][1]
    w0, rm0 = bn.weight.detach().clone(), bn.running_mean.clone()
    for i in range(3):
        b = {k: v.cuda() for k, v in ostep.make_batch(B, seed=20 + i).items()}
        m = tr.step(b)
        assert all(math.isfinite(v) for v in m.values()), m
    tr.sync_buffers()
    assert int(bn.num_batches_tracked) == 4
    assert not torch.equal(bn.weight.detach(), w0) and not torch.equal(bn.running_mean, rm0)
    assert torch.isfinite(tr.flat_g.data).all()
