"""GPU parity of the frozen identity network (restated ResNet18-128 / FeatureExtractModel) and of the identity-preserving
loss against oracle/identity_port.py (fp32 PyTorch on the CPU).

Parity here is ORACLE-DEFINED ("unpinned"): the reference's ResNet.py / FeatureExtract.py cannot be constructed
(SURVEY.md 2.3), so there is no reference output to record; the oracle restates what those files describe.

Tolerances (||a-b||_2/||b||_2): features of the production TF32 path <= 3e-3 (21 stacked convolutions with folded
BatchNorm); the input gradient is compared WITHOUT activation-mask injection, so ReLU sign flips caused by the forward
deviation (17 ReLUs and a max-pool arg-max between the features and the image) show up as a sqrt(eps)-sized deviation
(see tests/test_model_gpu.py): measured 8e-2 for TF32 (asserted 1.2e-1), <= 2e-2 in the fp32-exact verification mode,
which is the check that pins the backward plumbing; pooling kernels are compared bit-exactly / at fp32 round-off."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _net(seed=0, fdim=256, classes=347):
    from tpgan_b200.FeatureExtract import FeatureExtractModel
    from tpgan_b200.ResNet import BasicBlock
    torch.manual_seed(seed)
    net = FeatureExtractModel("resnet", classes, residualBlock=BasicBlock, feature_layer_dim_before_FC=fdim)
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():          # non-trivial BatchNorm statistics / affine parameters (a "pretrained" state)
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    net.eval()
    sd = {k: v.clone() for k, v in net.base_model.state_dict().items()}
    return net, sd


def test_pooling_kernels():
    from tpgan_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(3, 20, 33, 31)
    x[0, :, 4:7, 4:7] = 1.5          # ties inside windows: the first maximum in row-major order wins (ATen rule)
    xa = ops.Act.empty(3, 33, 31, 20).from_nchw(x.cuda())
    y = ops.Act.empty(3, 17, 16, 20)
    arg = torch.empty((3, 17, 16, 20), dtype=torch.uint8, device="cuda")
    ops.maxpool3s2(xa, y, arg)
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y.to_nchw().cpu(), ref.detach())
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dya = ops.Act.empty(3, 17, 16, 20).from_nchw(dy.cuda())
    dx = ops.Act.empty(3, 33, 31, 20)
    ops.maxpool3s2_backward(dya, arg, dx)
    assert rel(dx.to_nchw(), xr.grad) < 1e-6
    p = ops.Act.empty(3, 1, 1, 20)
    ops.avgpool(xa, p)
    assert rel(p.to_nchw().flatten(1), x.mean((2, 3))) < 1e-6
    g = torch.randn(3, 20)
    ga = ops.Act.empty(3, 1, 1, 20).from_nchw(g.view(3, 20, 1, 1).cuda())
    ops.avgpool_backward(ga, dx)
    assert rel(dx.to_nchw(), (g / (33 * 31)).view(3, 20, 1, 1).expand(3, 20, 33, 31)) < 1e-6


def test_identity_network_forward():
    from oracle import identity_port as ip
    net, sd = _net()
    x = torch.rand(4, 3, 128, 128, generator=torch.Generator().manual_seed(5)) * 2 - 1
    out, fc0 = net.cuda()(x.cuda())
    ref_out, ref_fc0, _ = ip.resnet18_128(sd, x)
    assert out.shape == (4, 347) and fc0.shape == (4, 256)
    assert rel(fc0, ref_fc0) < 3e-3, rel(fc0, ref_fc0)
    assert rel(out, ref_out) < 3e-3, rel(out, ref_out)
    with pytest.raises(RuntimeError):
        net(x)                       # CPU tensors: no fallback
    net.train()                      # training mode is the unfolded batch-statistics path (test_resnet_training_mode_vs_oracle)
    out_t, fc0_t = net(x.cuda())
    assert out_t.shape == (4, 347) and fc0_t.shape == (4, 256) and out_t.requires_grad


@pytest.mark.parametrize("exact", [False, True])
def test_identity_loss_and_input_gradient(exact):
    """Loss value against the oracle; the backward chain (FC0 -> pool -> 8 residual blocks -> max-pool -> stem, input
    gradient ACCUMULATED into d fake) is checked with fixed feature cotangents instead of the L1 sign pattern, which a
    forward deviation of 1e-3 flips for every feature whose fake/gt difference is that small."""
    from oracle import identity_port as ip
    from tpgan_b200 import ops
    from tpgan_b200.train_step import IdentityPlan
    net, sd = _net(seed=3)
    net.cuda()
    B = 3
    g = torch.Generator().manual_seed(9)
    fake = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).requires_grad_(True)
    gt = torch.rand(B, 3, 128, 128, generator=g) * 2 - 1
    c_pool, c_fc0 = torch.randn(B, 512, generator=g), torch.randn(B, 256, generator=g)
    ref = ip.identity_loss(sd, fake, gt)
    _, f0, p0 = ip.resnet18_128(sd, fake)
    ((p0 * c_pool).sum() + (f0 * c_fc0).sum()).backward()
    fa = ops.Act.empty(B, 128, 128, 3).from_nchw(fake.detach().cuda(), round_tf32=not exact)
    ga = ops.Act.empty(B, 128, 128, 3).from_nchw(gt.cuda(), round_tf32=not exact)
    dfa = ops.Act.empty(B, 128, 128, 3)
    base = torch.full((B, 3, 128, 128), 0.25)
    sums = torch.zeros(2, device="cuda")
    plan = IdentityPlan(net, B, fa, dfa, ga, 1.0, sums, torch.device("cuda"), exact=exact)
    # (1) the loss as scheduled
    dfa.from_nchw(base.cuda())
    for f in plan.schedule():
        f()
    torch.cuda.synchronize()
    got = plan.value(sums.cpu().tolist())
    assert abs(got - float(ref.detach())) <= (3e-4 if exact else 5e-3) * abs(float(ref.detach())), (got, float(ref.detach()))
    assert float((dfa.to_nchw().cpu() - base).abs().max()) > 0
    # (2) fixed cotangents through the same backward launches
    dfa.from_nchw(base.cuda())
    seeds = [plan.pf.grad_act(t) for t in plan.feats]
    for f in plan.pg.fwd + plan.pf.fwd:
        f()
    seeds[0].from_nchw(c_pool.view(B, 512, 1, 1).cuda())
    seeds[1].from_nchw(c_fc0.view(B, 256, 1, 1).cuda())
    for f in plan.pf.bwd:
        f()
    torch.cuda.synchronize()
    e = rel(dfa.to_nchw().cpu() - base, fake.grad)
    assert e < (2e-2 if exact else 1.2e-1), e


def test_step_with_identity_loss():
    """The fused G+D step with the identity term: every metric against the oracle step (same weights, same batch)."""
    from oracle import step as ostep
    from tpgan_b200 import D_and_G_model as M, config
    from tpgan_b200.train_step import TPGANTrainer
    M.EXACT_MODE = False
    torch.manual_seed(0)
    G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"])
    D = M.Discriminator(config.D["use_batchnorm"])
    pg = {k: v.clone().requires_grad_(True) for k, v in G.state_dict().items()}
    pd = {k: v.clone().requires_grad_(True) for k, v in D.state_dict().items()}
    net, sd = _net(seed=7)
    B = 2
    b = ostep.make_batch(B)
    tr = TPGANTrainer(G.cuda(), D.cuda(), B, identity_net=net.cuda())
    m = tr.step({k: v.cuda() for k, v in b.items()}, optimize=False)
    Gc, Dc = ostep.port_callables(pg, pd)
    og = torch.optim.Adam(list(pg.values()), lr=ostep.LEARNING_RATE)
    od = torch.optim.Adam(list(pd.values()), lr=ostep.LEARNING_RATE)
    ref = ostep.train_step(Gc, Dc, list(pg.values()), list(pd.values()), og, od, b, step_optim=False, identity_sd=sd)
    for k in ("pixel", "local", "symmetry", "tv", "ce", "ip", "g_total"):
        assert abs(m[k] - ref[k]) <= 1e-2 * abs(ref[k]) + 1e-4, (k, m[k], ref[k])


@pytest.mark.parametrize("exact,use_bn", [(True, True), (False, True), (True, False)])
def test_resnet_training_mode_vs_oracle(exact, use_bn):
    """Pre-training mode of the feature extractor (BASELINE config 5, "ResNet backbones"): unfolded conv -> batch-statistics
    BatchNorm (+ReLU, + shortcut add) forward and backward through the autograd bridge, against the oracle port in training
    mode.  fp32-exact mode: outputs <= 2e-4, running statistics <= 1e-4, parameter gradients <= 2e-2 overall (ReLU / max-pool
    gate flips, no mask injection); TF32: outputs <= 1e-2, gradients sanity-bounded (0.3)."""
    import math

    import tpgan_b200.D_and_G_model as M
    from oracle import identity_port as ip
    from tpgan_b200.ResNet import BasicBlock, ResNet18
    M.EXACT_MODE = exact
    try:
        torch.manual_seed(3)
        net = ResNet18(BasicBlock, 347, use_bn, 256)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        params = {k: sd[k].requires_grad_(True) for k, _ in net.named_parameters()}
        g = torch.Generator().manual_seed(4)
        x = torch.rand((6, 3, 128, 128), generator=g) * 2 - 1
        labels = torch.randint(0, 347, (6,), generator=g)
        lo, f0, _ = ip.resnet18_128(sd, x, training=True)
        loss = F.cross_entropy(lo, labels) + 0.1 * f0.square().mean()
        loss.backward()
        net.cuda().train()
        lg, fg = net(x.cuda())
        tol = 2e-4 if exact else 1e-2
        assert rel(lg, lo) < tol and rel(fg, f0) < tol, (rel(lg, lo), rel(fg, f0))
        lossg = F.cross_entropy(lg, labels.cuda()) + 0.1 * fg.square().mean()
        lossg.backward()
        if use_bn:
            sg = net.state_dict()
            for k in sd:
                if "running_" in k:
                    assert rel(sg[k], sd[k]) < (1e-4 if exact else 5e-3), k
            assert int(sg["conv1.1.num_batches_tracked"]) == 1
        num = den = 0.0
        worst = (0.0, "")
        for k, p in net.named_parameters():
            a, b = p.grad.double().cpu(), params[k].grad.double()
            num += float((a - b).pow(2).sum())
            den += float(b.pow(2).sum())
            r = float((a - b).norm() / (b.norm() + 1e-30))
            worst = max(worst, (r, k))
        overall = math.sqrt(num / den)
        assert overall < (2e-2 if exact else 0.3), (overall, worst)
        # back to eval: the folded copies are rebuilt from the trained weights / updated statistics
        net.eval()
        with torch.no_grad():
            le, fe = net(x.cuda())
        want_l, want_f, _ = ip.resnet18_128({k: v.detach() for k, v in sd.items()}, x, training=False)
        assert rel(le, want_l) < (1e-3 if exact else 1e-2) and rel(fe, want_f) < (1e-3 if exact else 1e-2)
    finally:
        M.EXACT_MODE = None


def test_classifier_trainer_step_vs_oracle():
    """The fused pre-training step of the ResNet backbone (ClassifierTrainer, fp32-exact mode, CUDA-graph replay from the
    second call): loss and SGD-Nesterov update against the oracle port driven by torch.optim.SGD, re-synchronised per step."""
    import math

    from oracle import identity_port as ip
    from oracle.pretrain_port import SGD
    from tpgan_b200.FeatureExtract import FeatureExtractModel
    from tpgan_b200.pretrain_step import ClassifierTrainer
    from tpgan_b200.ResNet import BasicBlock
    torch.manual_seed(5)
    model = FeatureExtractModel("resnet", 347, residualBlock=BasicBlock, feature_layer_dim_before_FC=256)
    net = model.base_model
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    params = [sd[k].requires_grad_(True) for k in names]
    opt = torch.optim.SGD(params, **SGD)
    model.cuda()
    B = 6
    tr = ClassifierTrainer(model, B, exact=True, use_graphs=True)
    pg = dict(net.named_parameters())
    for it in range(3):
        g = torch.Generator().manual_seed(30 + it)
        x, y = torch.rand((B, 3, 128, 128), generator=g) * 2 - 1, torch.randint(0, 347, (B,), generator=g)
        before = {k: sd[k].detach().clone() for k in names}
        opt.zero_grad()
        want = F.cross_entropy(ip.resnet18_128(sd, x, training=True)[0], y)
        want.backward()
        opt.step()
        m = tr.step(x.cuda(), y.cuda())
        assert abs(m["loss"] - float(want)) <= 1e-3 * abs(float(want)), (it, m, float(want))
        num = den = 0.0
        for k in names:
            d_g, d_w = (pg[k].detach().cpu() - before[k]).double(), (sd[k].detach() - before[k]).double()
            num += float((d_g - d_w).pow(2).sum())
            den += float(d_w.pow(2).sum())
        assert math.sqrt(num / den) < 5e-2, (it, math.sqrt(num / den))
        for k, p in zip(names, params):          # re-synchronise parameters + momentum, re-pack the tensor-core copies
            pg[k].data.copy_(p.detach())
            o = tr.flat.offsets["base_model." + k]
            tr.flat.m[o:o + p.numel()].copy_(opt.state[p]["momentum_buffer"].flatten())
        for L in tr.plan.layers:
            L.repack()


def test_eval_forward_follows_weight_updates():
    """the eval-mode network caches its BatchNorm-folded tensor-core layers.  A weight update between two eval
    forwards (optimizer step through torch, a train-mode forward that moves the running statistics, load_state_dict) must
    show up in the second one - the cache is keyed on parameter / buffer versions and re-folds in place."""
    from oracle import identity_port as ip
    net, _ = _net(seed=11)
    net = net.cuda()
    base = net.base_model
    g = torch.Generator().manual_seed(5)
    x = torch.rand((3, 3, 128, 128), generator=g) * 2 - 1
    xc = x.cuda()
    with torch.no_grad():
        f1, _ = base(xc)
    # (1) an optimizer-style in-place update of every parameter
    with torch.no_grad():
        for p in base.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g).cuda() * p.abs().mean())
        f2, z2 = base(xc)
    sd = {k: v.detach().cpu().clone() for k, v in base.state_dict().items()}
    want, want0, _ = ip.resnet18_128(sd, x, training=False)
    assert rel(f2, want) < 1e-2 and rel(z2, want0) < 1e-2, (rel(f2, want), rel(z2, want0))
    assert rel(f2, f1) > 1e-2            # the update is visible at all
    # (2) a train-mode forward moves the running statistics through raw device pointers
    base.train()
    base(xc)
    base.eval()
    with torch.no_grad():
        f3, _ = base(xc)
    sd = {k: v.detach().cpu().clone() for k, v in base.state_dict().items()}
    want3, _, _ = ip.resnet18_128(sd, x, training=False)
    assert rel(f3, want3) < 1e-2 and rel(f3, f2) > 1e-3, (rel(f3, want3), rel(f3, f2))
    # (3) load_state_dict back to the state after (1)
    sd1 = {k: v.clone() for k, v in base.state_dict().items()}
    base.load_state_dict({k: (v * 0.9 if v.is_floating_point() else v) for k, v in sd1.items()})
    with torch.no_grad():
        f4, _ = base(xc)
    want4, _, _ = ip.resnet18_128({k: v.detach().cpu() for k, v in base.state_dict().items()}, x, training=False)
    assert rel(f4, want4) < 1e-2, rel(f4, want4)
    with pytest.raises(RuntimeError, match="no autograd graph"):
        base(xc.clone().requires_grad_(True))
