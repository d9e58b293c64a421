"""GPU parity of the Pretrain path (MobileNetV2 + SSDHead + MultiTaskLoss + SGD-Nesterov step) against
oracle/pretrain_port.py (fp32 PyTorch on the CPU, pinned bit-exact to the live reference and to its Temp.py known answer by
tests/test_pretrain_cpu.py).

Tolerances, ||a-b||_2/||b||_2 unless stated:
  * depthwise conv / BatchNorm / SGD kernels (fp32 CUDA-core arithmetic): <= 2e-5 (summation order only);
  * MultiTaskLoss: assignment labels BIT-EXACT, loss and gradients <= 2e-5;
  * network outputs: fp32-exact verification mode (3xTF32 split) <= 2e-4 against the fp32 oracle; the production TF32 path
    in eval mode <= 1e-3 against the oracle WITH THE SAME TF32 ROUNDING POINTS (oracle.pretrain_port.forward_tf32_emulated
    - the sharp check of the TF32 path) and <= 2e-2 against the fp32 oracle; in training mode <= 3e-2 / 8e-2.  The last number is the network, not the kernels: at random initialisation, with
    batch-statistics BatchNorm over 64-256 samples per channel in the 4x4 / 8x8 stages, MobileNetV2 amplifies ANY
    perturbation ~100x from the stem to the heads (measured with tools/dbg_mbv2.py: fp32-exact mode 4e-7 -> 2.7e-5, TF32
    3.5e-4 -> 5e-2, every conv within 4e-4 of the oracle given the oracle's input - tests/test_conv_gpu.py);
  * parameter gradients: fp32-exact mode <= 2e-2 overall; TF32 against the fp32 oracle only sanity-bounded (0.6, measured
    0.42: the forward deviation above through the gates, amplified again by the backward pass).  The TF32 backward is
    pinned piecewise instead: every tensor-core dgrad / wgrad against ATen at <= 1e-3 and by adjoint identities
    (tests/test_conv_gpu.py), the fp32 kernels here at round-off, and the composition - the SAME traced plan, only the
    rounding flags and the 3x split differ - by the exact mode.  Gradients are compared WITHOUT
    activation-mask injection: a ReLU6 / ReLU gate that flips between the two implementations changes that element's
    gradient by 100 %, so a forward deviation eps shows up as a ~sqrt(eps) gradient deviation (tests/test_model_gpu.py
    discusses this); the exact mode is the check that pins the backward plumbing.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _act(t):
    from tpgan_b200 import ops
    n, c, h, w = t.shape
    return ops.Act.empty(n, h, w, c).from_nchw(t.cuda())


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("C,H,W,stride", [(32, 64, 64, 1), (96, 33, 17, 2), (144, 16, 16, 2), (960, 4, 4, 1), (8, 1, 1, 1),
                                          (24, 2, 2, 2)])
def test_depthwise_conv_kernels(C, H, W, stride):
    from tpgan_b200 import ops
    torch.manual_seed(C + H)
    x = torch.randn(3, C, H, W)
    w = torch.randn(C, 1, 3, 3) * 0.3
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride, 1, 1, C)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    Ho, Wo = ref.shape[2], ref.shape[3]
    xa, wa = _act(x), w.cuda()
    y = ops.Act.empty(3, Ho, Wo, C)
    ops.dwconv3x3(xa, y, wa, stride)
    assert rel(y.to_nchw(), ref.detach()) < 2e-6
    dya = _act(dy)
    dx = ops.Act.empty(3, H, W, C)
    ops.dwconv3x3_dgrad(dya, dx, wa, stride, False)
    assert rel(dx.to_nchw(), xr.grad) < 2e-6
    ops.dwconv3x3_dgrad(dya, dx, wa, stride, True)          # accumulate: 2x
    assert rel(dx.to_nchw(), 2 * xr.grad) < 2e-6
    dw = torch.zeros_like(wa)
    ops.dwconv3x3_wgrad(xa, dya, dw, stride)
    assert rel(dw, wr.grad) < 2e-5


@pytest.mark.parametrize("C,M,relu6,residual", [(32, (4, 64, 64), True, False), (16, (3, 9, 7), False, True),
                                                 (1280, (2, 4, 4), True, False), (96, (5, 1, 1), False, False)])
def test_batchnorm_kernels(C, M, relu6, residual):
    from tpgan_b200 import ops
    n, h, w = M
    torch.manual_seed(C)
    x = torch.randn(n, C, h, w) * 1.7 + 0.8
    r = torch.randn(n, C, h, w) if residual else None
    bn = torch.nn.BatchNorm2d(C)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.5 if not relu6 else 2.0)      # relu6: push a good share of the outputs beyond 6 / below 0
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    for training in (True, False):
        bn.train(training)
        bn.running_mean.copy_(rm0), bn.running_var.copy_(rv0)
        xr = x.clone().requires_grad_(True)
        rr = r.clone().requires_grad_(True) if residual else None
        o = bn(xr)
        if residual:
            o = o + rr
        if relu6:
            o = F.relu6(o)
        dy = torch.randn_like(o)
        bn.zero_grad()
        o.backward(dy)
        g, b = bn.weight.data.cuda(), bn.bias.data.cuda()
        rm, rv = rm0.cuda(), rv0.cuda()
        sums = torch.zeros(2 * C + 1, dtype=torch.float64, device="cuda")
        dsums = torch.zeros_like(sums)
        coef = torch.zeros(4 * C, device="cuda")
        xa, y = _act(x), ops.Act.empty(n, h, w, C)
        ops.bn_forward(xa, _act(r) if residual else None, y, g, b, rm, rv, 0.1, bn.eps, training, relu6, False, sums, coef)
        assert rel(y.to_nchw(), o.detach()) < 2e-5, training
        if n * h * w > 1:
            assert rel(rm, bn.running_mean) < 1e-5 and rel(rv, bn.running_var) < 1e-5
        dx = ops.Act.empty(n, h, w, C)
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        ops.bn_backward(_act(dy), xa, dx, coef, training, relu6, False, False, dsums, dg, db)
        tol = 5e-5 if n * h * w > 8 else 2e-3     # tiny batches: 1/sqrt(var) amplifies round-off in both implementations
        assert rel(dx.to_nchw(), xr.grad) < tol, training
        assert rel(dg, bn.weight.grad) < tol and rel(db, bn.bias.grad) < tol       # eval mode too (frozen statistics)
        assert float(sums.abs().sum()) == 0.0 and float(dsums.abs().sum()) == 0.0     # scratch left zeroed (incl. the ticket)


def test_rows_gather_round_trip():
    from tpgan_b200 import ops
    torch.manual_seed(0)
    a, b = torch.randn(3, 30, 4, 4), torch.randn(3, 12, 2, 2)      # 30 channels: padded to 32 in the NHWC buffer
    flat = torch.zeros(3, 532, device="cuda")
    ops.rows_gather(_act(a), flat, 532, 0)
    ops.rows_gather(_act(b), flat, 532, 480)
    want = torch.cat([a.permute(0, 2, 3, 1).reshape(3, -1), b.permute(0, 2, 3, 1).reshape(3, -1)], 1)
    assert torch.equal(flat[:, :528].cpu(), want)
    back = ops.Act.empty(3, 4, 4, 30)
    ops.rows_gather(back, flat, 532, 0, reverse=True)
    assert torch.equal(back.to_nchw().cpu(), a)


def _loss_case(B, n, seed, spread=140.0):
    g = torch.Generator().manual_seed(seed)
    loc = (torch.rand((B, n, 2), generator=g) * spread - 6).clamp_min(0)
    cls = torch.randn((B, n, 5), generator=g)
    true = torch.tensor([39.48, 40.28, 85.96, 38.7, 63.64, 63.65, 64.78, 89.32]) + torch.rand((B, 8), generator=g) * 6 - 3
    u = torch.rand((B, n), generator=g)
    return loc, cls, true, u


@pytest.mark.parametrize("B,n,seed", [(1, 394, 0), (8, 394, 1), (3, 50, 2), (2, 1000, 3)])
def test_multitask_loss_matches_oracle(B, n, seed):
    from oracle import pretrain_port as P
    from tpgan_b200.MobileNetV2 import MultiTaskLoss
    loc, cls, true, u = _loss_case(B, n, seed)
    if seed == 2:
        loc[:, :5] = loc[:, 5:10]            # exact ties between points
        u[:, 3] = u[:, 4]                    # and between keys
    lo, co = loc.clone().requires_grad_(True), cls.clone().requires_grad_(True)
    labs = []
    want = P.multitask_loss(lo, co, true, (128, 128), u, labels_out=labs)
    want.backward()
    L = MultiTaskLoss()
    lg, cg = loc.cuda().requires_grad_(True), cls.cuda().requires_grad_(True)
    got = L(lg, cg, true.cuda(), (128, 128), u.cuda())
    got.backward()
    assert torch.equal(L.labels.cpu(), torch.stack(labs)), "assignment must be bit-exact"
    assert abs(float(got) - float(want)) <= 2e-5 * abs(float(want))
    assert rel(lg.grad, lo.grad) < 2e-5 and rel(cg.grad, co.grad) < 2e-5


def test_multitask_loss_known_answer():
    """Temp.py:8-29 of the reference -> 0.8939134478569031 (its only published vector)."""
    from tpgan_b200.MobileNetV2 import MultiTaskLoss
    loc = torch.tensor([[[1.0, 1.0], [420.0, 360.0], [370.0, 150.0], [180.0, 220.0], [330.0, 270.0], [290.0, 135.0],
                         [500.0, 380.0], [190.0, 400.0], [210.0, 420.0], [510.0, 70.0], [178.0, 321.0], [420.0, 110.0]]])
    true = torch.tensor([[0.0, 0.0, 150.0, 400.0, 350.0, 250.0, 300.0, 150.0]])
    cls = torch.tensor([[[2.0, 1.0, 0.1, 0.5, 1.4], [1.0, 2.0, 0.1, 0.3, 1.1], [0.1, 2.0, 1.0, 0.4, 0.5],
                         [2.0, 0.1, 1.0, 0.7, 0.5], [1.0, 0.1, 1.4, 0.8, 2.0], [0.1, 1.0, 2.0, 0.6, 0.7],
                         [2.0, 1.0, 0.1, 0.9, 1.5], [1.0, 0.8, 0.1, 1.1, 2.0], [0.1, 1.2, 1.0, 2.0, 0.5],
                         [2.0, 0.1, 1.0, 1.3, 0.6], [1.0, 0.1, 2.0, 1.4, 1.6], [0.1, 1.0, 1.3, 1.5, 2.0]]])
    got = MultiTaskLoss()(loc.cuda(), cls.cuda(), true.cuda(), (600, 800))
    assert abs(float(got) - 0.8939134478569031) < 2e-6


def test_sgd_nesterov_matches_torch():
    from oracle.pretrain_port import SGD
    from tpgan_b200 import ops
    torch.manual_seed(0)
    p0 = torch.randn(4096)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref], **SGD)
    p, buf = p0.clone().cuda(), torch.zeros(4096, device="cuda")
    lr = torch.full((1,), SGD["lr"], device="cuda")
    for _ in range(4):
        g = torch.randn(4096)
        ref.grad = g.clone()
        opt.step()
        ops.sgd_step(p, g.cuda(), buf, lr, SGD["momentum"], SGD["weight_decay"], SGD["nesterov"])
    assert rel(p, ref.data) < 1e-6 and rel(p - p0.cuda(), ref.data - p0) < 1e-4


# ------------------------------------------------------------------------------------------------ network
def _nets(seed=0, randomize_bn=True):
    from oracle.pretrain_port import MobileNetV2Port
    from tpgan_b200.MobileNetV2 import MobileNetV2
    torch.manual_seed(seed)
    port = MobileNetV2Port()
    if randomize_bn:   # non-trivial affine parameters / running statistics and non-zero head biases
        g = torch.Generator().manual_seed(seed + 1)
        for m in port.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            elif isinstance(m, torch.nn.Conv2d) and m.bias is not None:
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    net = MobileNetV2()
    net.load_state_dict(port.state_dict())
    return port, net.cuda()


def _set_exact(flag):
    import tpgan_b200.D_and_G_model as M
    M.EXACT_MODE = flag


@pytest.mark.parametrize("exact", [False, True])
def test_network_forward_backward_vs_oracle(exact):
    from oracle.pretrain_port import make_batch
    _set_exact(exact)
    try:
        port, net = _nets()
        sd0 = {k: v.clone() for k, v in port.state_dict().items()}
        x, _, _ = make_batch(4, seed=7)
        port.train(), net.train()
        lw, cw = port(x)
        lg, cg = net(x.cuda())
        tol = 2e-4 if exact else 8e-2
        assert lg.shape == (4, 394, 2) and cg.shape == (4, 394, 5)
        assert rel(lg, lw) < tol and rel(cg, cw) < tol, (rel(lg, lw), rel(cg, cw))
        if not exact:      # the sharp check of the TF32 path: same rounding points in the oracle
            from oracle.pretrain_port import MobileNetV2Port, forward_tf32_emulated
            twin = MobileNetV2Port()
            twin.load_state_dict(sd0)
            twin.train()
            le, ce = forward_tf32_emulated(twin, x)
            assert rel(lg, le) < 3e-2 and rel(cg, ce) < 3e-2, (rel(lg, le), rel(cg, ce))
        # running statistics after one training forward
        sw, sg = port.state_dict(), net.state_dict()
        for k in sw:
            if "running_" in k:
                assert rel(sg[k], sw[k]) < (1e-4 if exact else 5e-2), k
            elif "num_batches_tracked" in k:
                assert int(sg[k]) == int(sw[k]) == 1
        # backward with the same upstream gradients
        gen = torch.Generator().manual_seed(3)
        dl, dc = torch.randn(lw.shape, generator=gen) * (lw > 0), torch.randn(cw.shape, generator=gen)
        for p in port.parameters():
            p.grad = None
        torch.autograd.backward([lw, cw], [dl, dc])
        torch.autograd.backward([lg, cg], [dl.cuda(), dc.cuda()])
        num = den = 0.0
        worst = (0.0, "")
        pg = dict(net.named_parameters())
        for k, p in port.named_parameters():
            a, b = pg[k].grad.double().cpu(), p.grad.double()
            num += float((a - b).pow(2).sum())
            den += float(b.pow(2).sum())
            r = float((a - b).norm() / (b.norm() + 1e-30))
            if r > worst[0]:
                worst = (r, k)
        overall = math.sqrt(num / den)
        assert overall < (2e-2 if exact else 0.6), (overall, worst)
    finally:
        _set_exact(None)


def test_network_eval_mode_uses_running_statistics():
    from oracle.pretrain_port import make_batch
    port, net = _nets(seed=2)
    x, _, _ = make_batch(3, seed=9)
    port.eval(), net.eval()
    with torch.no_grad():
        lw, cw = port(x)
        lg, cg = net(x.cuda())
    assert rel(lg, lw) < 2e-2 and rel(cg, cw) < 2e-2, (rel(lg, lw), rel(cg, cw))
    from oracle.pretrain_port import forward_tf32_emulated
    le, ce = forward_tf32_emulated(port, x)
    assert rel(lg, le) < 1e-3 and rel(cg, ce) < 1e-3, (rel(lg, le), rel(cg, ce))
    sw, sg = port.state_dict(), net.state_dict()
    assert all(torch.equal(sg[k].cpu(), sw[k]) for k in sw if "running_" in k)   # untouched in eval mode


def test_module_rejects_cpu_tensors():
    from tpgan_b200.MobileNetV2 import MobileNetV2, MultiTaskLoss
    with pytest.raises(RuntimeError):
        MobileNetV2()(torch.zeros(1, 3, 128, 128))
    with pytest.raises(RuntimeError):
        MultiTaskLoss()(torch.zeros(1, 12, 2), torch.zeros(1, 12, 5), torch.zeros(1, 8), (128, 128))


# ------------------------------------------------------------------------------------------------ training step
@pytest.mark.parametrize("graphs,exact", [(False, True), (True, False)])
def test_pretrain_step_vs_oracle(graphs, exact):
    """Three optimisation steps against the oracle step driven by torch.optim.SGD.  Every step starts from IDENTICAL state
    (parameters and momentum buffers are re-synchronised from the oracle after each comparison): with alpha = 30 and the
    reference's learning rate the first updates change some weights by >10 %, and the network amplifies a 1 % difference
    of such an update into a different trajectory within two steps - in any implementation.  Per step: loss, assignment
    and parameter UPDATE.  fp32-exact mode pins the step: loss <= 1e-3, labels > 99 % equal, update <= 0.2 overall (measured
    < 5e-2 in steps 0-1, 0.11 in step 2: the oracle's loss is evaluated on ITS OWN predictions; ReLU6 gate flips, see
    module docstring).  The TF32 graph-replayed production mode is only SANITY-bounded end to end: loss <= 0.12 (measured 2-6 %), labels
    > 90 % equal, update <= 1.2 (measured 0.88: ~10 % of the points get another label - the assignment is discontinuous
    in the predictions - so the loss gradients themselves differ by ~30 % before the backward pass amplifies them).  In
    BOTH modes the loss kernel is bit-exact in its assignment once the oracle loss is fed the CUDA path's own predictions."""
    from oracle import pretrain_port as P
    from tpgan_b200.pretrain_step import PretrainTrainer
    port, net = _nets(seed=4)
    B = 4
    opt = torch.optim.SGD(port.parameters(), **P.SGD)
    tr = PretrainTrainer(net, B, use_graphs=graphs, exact=exact)
    ltol, atol, utol = (1e-3, 0.99, 0.2) if exact else (0.12, 0.90, 1.2)
    pw, pg = dict(port.named_parameters()), dict(net.named_parameters())
    for it in range(3):
        before = {k: v.detach().clone() for k, v in pw.items()}
        x, true, u = P.make_batch(B, seed=20 + it)
        want, labs, _, _ = P.pretrain_step(port, x, true, u, opt)
        m = tr.step(x.cuda(), true.cuda(), u.cuda())
        assert abs(m["loss"] - float(want)) <= ltol * abs(float(want)), (it, m, float(want))
        agree = float((tr.labels.cpu() == torch.stack(labs)).float().mean())
        assert agree > atol, (it, agree)
        lg, cg = tr.outputs()
        own = []
        own_loss = P.multitask_loss(lg.cpu(), cg.cpu(), true, (128, 128), u, labels_out=own)
        assert torch.equal(tr.labels.cpu(), torch.stack(own)), it
        assert abs(m["loss"] - float(own_loss)) <= 2e-5 * abs(float(own_loss)), (it, m, float(own_loss))
        num = den = 0.0
        for k, p in pg.items():
            d_g, d_w = (p.detach().cpu() - before[k]).double(), (pw[k].detach() - before[k]).double()
            num += float((d_g - d_w).pow(2).sum())
            den += float(d_w.pow(2).sum())
        assert math.sqrt(num / den) < utol, (it, math.sqrt(num / den))
        for k, p in pg.items():          # re-synchronise parameters and momentum, then re-pack the tensor-core copies
            p.data.copy_(pw[k].detach())
            o = tr.flat.offsets[k]
            tr.flat.m[o:o + p.numel()].copy_(opt.state[pw[k]]["momentum_buffer"].flatten())
        for L in tr.plan.layers:
            L.repack()
    tr.sync_buffers()
    assert int(net.conv1[1].num_batches_tracked) == 3


def test_full_size_step_properties():
    """B = 32 (the per-GPU batch of BASELINE config 5 at 8 GPUs): size-independent properties."""
    from oracle import pretrain_port as P
    from tpgan_b200.pretrain_step import PretrainTrainer
    _, net = _nets(seed=5, randomize_bn=False)
    B = 32
    tr = PretrainTrainer(net, B, use_graphs=True)
    x, true, u = P.make_batch(B, seed=40)
    xs, ts, us = x.cuda(), true.cuda(), u.cuda()
    m0 = tr.step(xs, ts, us, optimize=False)
    g0 = tr.flat.grad.clone()
    m1 = tr.step(xs, ts, us, optimize=False)
    # idempotent without the optimizer, up to the running statistics (not used in training mode) and atomics order
    assert abs(m0["loss"] - m1["loss"]) < 1e-5 * abs(m0["loss"]) and rel(tr.flat.grad, g0) < 1e-4
    lab = tr.labels.cpu()
    k = int(0.1 * tr.n)
    assert ((lab >= 0).sum(1) >= k).all() and ((lab >= 0).sum(1) <= 4 * k + 8).all()
    assert abs(m0["loss"] - (30.0 * m0["location"] + 0.1 * m0["classification"])) < 1e-4 * abs(m0["loss"])
    # BatchNorm in training mode: the stem's output has the batch statistics its affine parameters prescribe
    y = tr.plan.named["conv1.1"].act.buf
    pre = tr.plan.named["conv1.0"].act.buf.view(-1, 32)
    assert torch.isfinite(y).all() and float(y.min()) >= 0.0 and float(y.max()) <= 6.0
    mean, var = pre.mean(0), pre.var(0, unbiased=False)
    st = net.conv1[1]._tc_aux._states[id(tr.plan)]
    assert rel(st.coef[64:96], mean) < 1e-4 and rel(st.coef[96:128], (var + 1e-5).rsqrt()) < 1e-4
    # training decreases the loss on a fixed batch
    first = tr.step(xs, ts, us)["loss"]
    for _ in range(10):
        last = tr.step(xs, ts, us)["loss"]
    assert last < first, (first, last)


def test_checkpoint_helpers_round_trip(tmp_path):
    """save_model / save_optimizer (UtilityMethods.py:58-103) on the trainer's model and fused optimizer state: the files
    load into the oracle network and a stock torch.optim.SGD, and back into a fresh trainer."""
    from oracle import pretrain_port as P
    from tpgan_b200.MobileNetV2 import MobileNetV2
    from tpgan_b200.pretrain_step import PretrainTrainer
    from tpgan_b200.UtilityMethods import save_model, save_optimizer
    port, net = _nets(seed=6)
    tr = PretrainTrainer(net, 2)
    x, true, u = P.make_batch(2, seed=50)
    for _ in range(2):
        tr.step(x.cuda(), true.cuda(), u.cuda())
    tr.sync_buffers()
    mpath = save_model(net, str(tmp_path / "run"), 3)
    opath = save_optimizer(tr.optimizer, net, str(tmp_path / "run"), 3)
    assert mpath.endswith("model_epoch_3.pth") and opath.endswith("optimizer_epoch_3.pth")
    sd = torch.load(mpath, map_location="cpu")
    port.load_state_dict(sd, strict=True)                       # reference-layout keys and shapes
    ck = torch.load(opath, map_location="cpu")
    assert set(ck) == {"optimizer", "model", "epoch"} and ck["epoch"] == 3
    opt = torch.optim.SGD(port.parameters(), **P.SGD)
    opt.load_state_dict(ck["optimizer"])                        # a stock optimizer accepts the fused state
    bufs = [opt.state[p]["momentum_buffer"] for p in port.parameters()]
    assert len(bufs) == 197 and all(float(b.abs().sum()) > 0 for b in bufs[:5])
    g = opt.param_groups[0]
    assert (g["lr"], g["momentum"], g["weight_decay"], g["nesterov"]) == (5e-4, 0.9, 5e-4, True)
    # and back: a fresh trainer resumes with identical parameters and momentum
    torch.manual_seed(123)
    net2 = MobileNetV2().cuda()
    net2.load_state_dict(torch.load(mpath))
    tr2 = PretrainTrainer(net2, 2)
    tr2.optimizer.load_state_dict(torch.load(opath)["optimizer"])
    assert torch.equal(tr2.flat.m, tr.flat.m) and torch.equal(tr2.flat.data, tr.flat.data)
    a = tr.step(x.cuda(), true.cuda(), u.cuda())
    b = tr2.step(x.cuda(), true.cuda(), u.cuda())
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(a["loss"])
    assert rel(tr2.flat.data, tr.flat.data) < 1e-6


# ------------------------------------------------------------------------------------------------ decoder (row f4)
@pytest.mark.parametrize("top_k,thr,nms", [(1, 0.5, 20.0), (3, 0.6, 15.0), (8, 0.3, 40.0)])
def test_decoder_matches_oracle(top_k, thr, nms):
    from oracle import pretrain_port as P
    from tpgan_b200.MobileNetV2 import MultiTaskDecoder
    g = torch.Generator().manual_seed(int(nms))
    B, n = 6, 394
    loc = torch.rand((B, n, 2), generator=g) * 128
    cls = torch.randn((B, n, 5), generator=g) * 3
    loc[:, 7] = loc[:, 3]                                  # duplicate points (distance 0 <= threshold: suppressed)
    true = torch.rand((B, 8), generator=g) * 128
    true[0, 0:2] = loc[0, 0]
    dec = MultiTaskDecoder(thr, top_k, nms)
    count, score, point, acc = dec.decode(loc.cuda(), cls.cuda(), true.cuda())
    for b in range(B):
        c_w, s_w, p_w = P.decode_sample(loc[b], cls[b], thr, top_k, nms)
        assert torch.equal(count[b].cpu(), c_w), b
        assert torch.equal(point[b].cpu(), p_w), b                      # selected points: bit-exact
        assert torch.allclose(score[b].cpu(), s_w, rtol=0, atol=2e-6), b
        assert abs(float(acc[b]) - P.accuracy_sample(c_w, p_w, true[b])) < 1e-6, b
    out = dec(loc.cuda(), cls.cuda())                                   # the reference's list-of-tuples structure
    assert len(out) == B and sum(len(o) for o in out) == int(count.sum())
    c0, s0, p0 = out[0][0]
    assert isinstance(c0, int) and s0.shape == () and p0.shape == (2,)


def test_decoder_known_answer():
    """Temp.py:7-29,55-61: one detection - class 1, confidence 0.5148, point (370, 150)."""
    from tpgan_b200.MobileNetV2 import MultiTaskDecoder
    loc = torch.tensor([[[1.0, 1.0], [420.0, 360.0], [370.0, 150.0], [180.0, 220.0], [330.0, 270.0], [290.0, 135.0],
                         [500.0, 380.0], [190.0, 400.0], [210.0, 420.0], [510.0, 70.0], [178.0, 321.0], [420.0, 110.0]]])
    cls = torch.tensor([[[2.0, 1.0, 0.1, 0.5, 1.4], [1.0, 2.0, 0.1, 0.3, 1.1], [0.1, 2.0, 1.0, 0.4, 0.5],
                         [2.0, 0.1, 1.0, 0.7, 0.5], [1.0, 0.1, 1.4, 0.8, 2.0], [0.1, 1.0, 2.0, 0.6, 0.7],
                         [2.0, 1.0, 0.1, 0.9, 1.5], [1.0, 0.8, 0.1, 1.1, 2.0], [0.1, 1.2, 1.0, 2.0, 0.5],
                         [2.0, 0.1, 1.0, 1.3, 0.6], [1.0, 0.1, 2.0, 1.4, 1.6], [0.1, 1.0, 1.3, 1.5, 2.0]]])
    out = MultiTaskDecoder(nms_distance_threshold=30)(loc.cuda(), cls.cuda())[0]
    assert len(out) == 1
    c, s, p = out[0]
    assert c == 1 and abs(float(s) - 0.5148) < 5e-5 and p.tolist() == [370.0, 150.0]


def test_pretrain_loop_runs_and_checkpoints(tmp_path):
    """tpgan_b200.Pretrain.main: the reference's loop shape (train step, decode + accuracy, validation in eval mode,
    MultiStepLR, save_model / save_optimizer per epoch) on synthetic batches."""
    from tpgan_b200 import Pretrain as PT
    hist = PT.main(["--epochs", "2", "--steps-per-epoch", "3", "--batch", "4", "--log-every", "2", "--val-steps", "1",
                    "--log-dir", str(tmp_path / "log")])
    assert len(hist) == 6 and all(math.isfinite(l) and 0.0 <= a <= 1.0 for l, a in hist)
    for e in (0, 1):
        assert (tmp_path / "log" / f"model_epoch_{e}.pth").exists() and (tmp_path / "log" / f"optimizer_epoch_{e}.pth").exists()
    sd = torch.load(tmp_path / "log" / "model_epoch_1.pth", map_location="cpu")
    assert int(sd["conv1.1.num_batches_tracked"]) == 6


def test_full_size_forward_tf32_vs_emulated_oracle():
    """B = 32 (BASELINE config 5's per-GPU batch): the production TF32 forward against the oracle with the same rounding
    points - eval mode <= 1e-3, training mode (batch statistics over 512+ samples per channel) <= 3e-2."""
    from oracle.pretrain_port import forward_tf32_emulated, make_batch
    port, net = _nets(seed=8)
    x, _, _ = make_batch(32, seed=60)
    for train, tol in ((False, 1e-3), (True, 3e-2)):
        port.train(train), net.train(train)
        sd = {k: v.clone() for k, v in port.state_dict().items()}
        le, ce = forward_tf32_emulated(port, x)
        port.load_state_dict(sd)          # the emulated training forward updated the running statistics
        with torch.no_grad():
            lg, cg = net(x.cuda())
        assert rel(lg, le) < tol and rel(cg, ce) < tol, (train, rel(lg, le), rel(cg, ce))


def test_network_rejects_unsupported_aspect_ratio():
    from tpgan_b200.MobileNetV2 import MobileNetV2
    net = MobileNetV2().cuda().eval()
    with pytest.raises(NotImplementedError, match="extent 1 in one dimension"):
        net(torch.zeros(1, 3, 128, 64, device="cuda"))


@pytest.mark.parametrize("hw,batch", [((64, 64), 3), ((256, 256), 1)])
def test_network_other_image_sizes(hw, batch):
    """The reference runs batch 1 'to handle images of different spatial size' (config.py:12): the traced network follows the
    input size (plans are cached per shape).  fp32-exact mode against the oracle, forward and parameter gradients; the
    number of predicted points follows MobileNetV2.num_points."""
    import tpgan_b200.D_and_G_model as M
    from oracle.pretrain_port import MobileNetV2Port
    from tpgan_b200.MobileNetV2 import MobileNetV2
    M.EXACT_MODE = True
    try:
        port, net = _nets(seed=9)
        g = torch.Generator().manual_seed(1)
        x = torch.rand((batch, 3) + hw, generator=g) * 2 - 1
        port.eval(), net.eval()          # running statistics: a batch of 1 has no usable batch statistics at 1x1 maps
        lw, cw = port(x)
        lg, cg = net(x.cuda())
        n = MobileNetV2.num_points(*hw)
        assert lg.shape == (batch, n, 2) and cg.shape == (batch, n, 5) and lw.shape == lg.shape
        assert rel(lg, lw) < 2e-4 and rel(cg, cw) < 2e-4, (rel(lg, lw), rel(cg, cw))
        gen = torch.Generator().manual_seed(2)
        dl, dc = torch.randn(lw.shape, generator=gen), torch.randn(cw.shape, generator=gen)
        torch.autograd.backward([lw, cw], [dl, dc])
        torch.autograd.backward([lg, cg], [dl.cuda(), dc.cuda()])
        pg = dict(net.named_parameters())
        num = den = 0.0
        for k, p in port.named_parameters():
            assert pg[k].grad is not None, k      # incl. the BatchNorm affine parameters (frozen statistics, trainable)
            num += float((pg[k].grad.double().cpu() - p.grad.double()).pow(2).sum())
            den += float(p.grad.double().pow(2).sum())
        assert math.sqrt(num / den) < 2e-2, math.sqrt(num / den)
    finally:
        M.EXACT_MODE = None


def test_prefetched_batches_give_the_same_steps():
    """step(batch=..., prefetch_next=...) (host->device copies of the next batch on a copy stream) == step(images, labels, u).
    Steps run without the optimizer so that they are independent: with updates, the summation-order noise of the atomics
    (1e-6) is amplified into visibly different trajectories within two steps (see test_pretrain_step_vs_oracle)."""
    from oracle import pretrain_port as P
    from tpgan_b200.pretrain_step import PretrainTrainer
    batches = [tuple(t.pin_memory() for t in P.make_batch(4, seed=70 + i)) for i in range(4)]
    runs = []
    for mode in ("direct", "prefetch"):
        _, net = _nets(seed=11)
        tr = PretrainTrainer(net, 4, use_graphs=True)
        out = []
        if mode == "prefetch":
            tr.prefetch(batches[0])
        for i, b in enumerate(batches):
            if mode == "direct":
                m = tr.step(*[t.cuda() for t in b], optimize=False)
            else:
                m = tr.step(None, None, batch=b, prefetch_next=batches[i + 1] if i + 1 < len(batches) else None, optimize=False)
            out.append((m, tr.flat.grad.clone(), tr.labels.clone()))
        runs.append(out)
    for (ma, ga, la), (mb, gb, lb) in zip(*runs):
        assert abs(ma["loss"] - mb["loss"]) <= 1e-5 * abs(ma["loss"]), (ma, mb)
        assert torch.equal(la, lb) and rel(gb, ga) < 1e-4


def test_kernels_do_not_write_outside_their_outputs():
    """(compute-sanitizer is closed on this pool.)  Every output of the Pretrain-path kernels sits between two guard slabs
    filled with a sentinel; after the launches the guards are untouched.  Odd sizes on purpose (partial strips / groups)."""
    from tpgan_b200 import ops
    S = 1234.5

    def guarded(n, h, w, c):
        big = torch.full((n + 2, h, w, c), S, device="cuda")
        return big, ops.Act(big[1:-1])

    def intact(big):
        return bool((big[0] == S).all()) and bool((big[-1] == S).all())

    torch.manual_seed(0)
    for C, H, W, s in ((24, 13, 11, 1), (96, 9, 7, 2), (960, 4, 4, 1)):
        x = ops.Act(torch.randn(3, H, W, C, device="cuda"))
        Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
        w = torch.randn(C, 1, 3, 3, device="cuda")
        by, y = guarded(3, Ho, Wo, C)
        ops.dwconv3x3(x, y, w, s)
        bx, dx = guarded(3, H, W, C)
        ops.dwconv3x3_dgrad(ops.Act(torch.randn(3, Ho, Wo, C, device="cuda")), dx, w, s, False)
        dwb = torch.full((C + 2, 1, 3, 3), S, device="cuda")
        dwb[1:-1] = 0
        ops.dwconv3x3_wgrad(x, ops.Act(torch.randn(3, Ho, Wo, C, device="cuda")), dwb[1:-1], s)
        assert intact(by) and intact(bx) and bool((dwb[0] == S).all()) and bool((dwb[-1] == S).all()), (C, H, W, s)
        # BatchNorm forward / backward
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        sums = torch.zeros(2 * C + 3, dtype=torch.float64, device="cuda")
        sums[-2:] = S
        dsums = sums.clone()
        coef = torch.full((4 * C + 8,), S, device="cuda")
        bo, o = guarded(3, H, W, C)
        ops.bn_forward(x, None, o, g, b, rm, rv, 0.1, 1e-5, True, True, False, sums[:2 * C + 1], coef[:4 * C])
        bd, d = guarded(3, H, W, C)
        dgb = torch.full((C + 8,), S, device="cuda")
        ops.bn_backward(ops.Act(torch.randn(3, H, W, C, device="cuda")), x, d, coef[:4 * C], True, True, False, False,
                        dsums[:2 * C + 1], dgb[:C], dgb[:C].clone())
        assert intact(bo) and intact(bd) and bool((coef[4 * C:] == S).all()) and bool((dgb[C:] == S).all())
        assert bool((sums[-2:] == S).all()) and bool((dsums[-2:] == S).all())
    # loss / decoder outputs
    B, n = 3, 77
    loc, cls = torch.rand(B, n, 2, device="cuda") * 128, torch.randn(B, n, 5, device="cuda")
    true, u = torch.rand(B, 8, device="cuda") * 128, torch.rand(B, n, device="cuda")
    dl, dc = torch.full((B + 2, n, 2), S, device="cuda"), torch.full((B + 2, n, 5), S, device="cuda")
    lab = torch.full((B + 2, n), 7, dtype=torch.int32, device="cuda")
    sums3 = torch.full((5,), S, device="cuda")
    sums3[:3] = 0
    ops.multitask_loss(loc, cls, true, u, n, 2 * n, 5 * n, 7, 128.0, 128.0, 30.0, 0.1, 5.0, 1.0 / B, dl[1:-1], dc[1:-1],
                       lab[1:-1], sums3[:3])
    assert intact(dl) and intact(dc) and bool((lab[0] == 7).all()) and bool((lab[-1] == 7).all()) and bool((sums3[3:] == S).all())
    cnt = torch.full((B + 2, 5), 9, dtype=torch.int32, device="cuda")
    sc, pt = torch.full((B + 2, 5, 3), S, device="cuda"), torch.full((B + 2, 5, 3, 2), S, device="cuda")
    ops.ssd_decode(loc, cls, n, 2 * n, 5 * n, 5, 3, 0.3, 10.0, cnt[1:-1], sc[1:-1], pt[1:-1])
    assert intact(sc) and intact(pt) and bool((cnt[0] == 9).all()) and bool((cnt[-1] == 9).all())
