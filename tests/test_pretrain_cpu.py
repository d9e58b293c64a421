"""CPU tests of the Pretrain path: the oracle (oracle/pretrain_port.py) is pinned against the live
reference when /root/reference is present and against the golden vectors recorded from it (tools/make_golden_pretrain.py ->
tests/golden/pretrain_golden.pt); the drop-in module mirrors the reference's interface and state_dict."""
import contextlib
import io
import os
import sys

import pytest
import torch

from oracle import pretrain_port as P

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pretrain_golden.pt")
REF = os.environ.get("TPGAN_REFERENCE_DIR", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "MobileNetV2.py")), reason="reference tree absent")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _ref():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import MobileNetV2 as R
    return R


def _temp_py_inputs():
    loc = torch.tensor([[[1.0, 1.0], [420.0, 360.0], [370.0, 150.0], [180.0, 220.0], [330.0, 270.0], [290.0, 135.0],
                         [500.0, 380.0], [190.0, 400.0], [210.0, 420.0], [510.0, 70.0], [178.0, 321.0], [420.0, 110.0]]])
    true = torch.tensor([[0.0, 0.0, 150.0, 400.0, 350.0, 250.0, 300.0, 150.0]])
    cls = torch.tensor([[[2.0, 1.0, 0.1, 0.5, 1.4], [1.0, 2.0, 0.1, 0.3, 1.1], [0.1, 2.0, 1.0, 0.4, 0.5],
                         [2.0, 0.1, 1.0, 0.7, 0.5], [1.0, 0.1, 1.4, 0.8, 2.0], [0.1, 1.0, 2.0, 0.6, 0.7],
                         [2.0, 1.0, 0.1, 0.9, 1.5], [1.0, 0.8, 0.1, 1.1, 2.0], [0.1, 1.2, 1.0, 2.0, 0.5],
                         [2.0, 0.1, 1.0, 1.3, 0.6], [1.0, 0.1, 2.0, 1.4, 1.6], [0.1, 1.0, 1.3, 1.5, 2.0]]])
    return loc, cls, true


def _loss_case(seed, n=394):     # the generator of tools/make_golden_pretrain.py
    g = torch.Generator().manual_seed(seed)
    loc = (torch.rand((1, n, 2), generator=g) * 140 - 6).clamp_min(0)
    cls = torch.randn((1, n, 5), generator=g)
    true = torch.tensor([[39.48, 40.28, 85.96, 38.7, 63.64, 63.65, 64.78, 89.32]]) + torch.rand((1, 8), generator=g) * 6 - 3
    u = torch.rand((1, n), generator=g)
    return loc, cls, true, u


def test_multitask_loss_known_answer(gold):
    """Temp.py:8-29, the reference's only known answer."""
    loc, cls, true = _temp_py_inputs()
    got = float(P.multitask_loss(loc, cls, true, (600, 800)))
    assert got == pytest.approx(0.8939134478569031, abs=1e-7)
    assert gold["temp_py_total_loss"] == 0.8939134478569031


def test_port_matches_golden(gold):
    torch.manual_seed(0)
    net = P.MobileNetV2Port()
    sums = {k: float(v.double().sum()) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    assert sums == gold["param_sums"]                       # same initialisation order / RNG consumption
    x, _, _ = P.make_batch(2, seed=11)
    net.train()
    with torch.no_grad():
        loc, cls = net(x)
    assert torch.equal(loc, gold["train_loc"]) and torch.equal(cls, gold["train_cls"])
    sd = net.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in gold["running"].items())
    net.eval()
    with torch.no_grad():
        loc, cls = net(x)
    assert torch.equal(loc, gold["eval_loc"]) and torch.equal(cls, gold["eval_cls"])


def test_loss_matches_golden(gold):
    for case in gold["loss_cases"]:
        loc, cls, true, u = _loss_case(case["seed"])
        loc.requires_grad_(True), cls.requires_grad_(True)
        labs = []
        val = P.multitask_loss(loc, cls, true, (128, 128), u, labels_out=labs)
        gl, gc = torch.autograd.grad(val, [loc, cls])
        assert torch.equal(labs[0], case["labels"])
        assert float(val) == pytest.approx(case["loss"], rel=1e-6)
        assert torch.allclose(gl, case["dloc"], rtol=1e-5, atol=1e-9) and torch.allclose(gc, case["dcls"], rtol=1e-5, atol=1e-9)


def test_assignment_rules():
    # threshold = k-th smallest distance with ties included; a point positive for two labels takes the nearer, first on ties
    d = torch.tensor([[1.0, 9.0, 9.0, 9.0], [2.0, 2.0, 9.0, 9.0], [2.0, 1.0, 9.0, 9.0], [5.0, 5.0, 1.0, 1.0]] +
                     [[7.0 + 0.1 * i] * 4 for i in range(16)])
    lab = P.assign_labels(d, ratio=0.1)            # n = 20, k = 2: thresholds 2, 2, 7, 7
    # point 1 ties label 0 / 1 at distance 2 -> first label; points 3 and 4 tie label 2 / 3 -> label 2
    assert lab[:5].tolist() == [0, 0, 1, 2, 2] and (lab[5:] == -1).all()
    # background sub-sampling keeps the m smallest keys, lower index on ties
    labels = torch.tensor([0, -1, -1, -1, -1, -1, -1, -1], dtype=torch.int32)
    u = torch.tensor([0.0, 0.9, 0.1, 0.5, 0.1, 0.7, 0.2, 0.3])
    sel = P.select_background(labels, u, ratio_nb=3.0)   # 1 positive -> 3 background points
    assert sel.nonzero().flatten().tolist() == [2, 4, 6]


@needs_ref
def test_port_matches_live_reference():
    R = _ref()
    torch.manual_seed(0)
    ref = R.MobileNetV2()
    torch.manual_seed(0)
    port = P.MobileNetV2Port()
    sr, sp = ref.state_dict(), port.state_dict()
    assert list(sr.keys()) == list(sp.keys()) and all(torch.equal(sr[k], sp[k]) for k in sr)
    x, _, _ = P.make_batch(2, seed=3)
    ref.train(), port.train()
    a, b = ref(x), port(x)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    (a[0].sum() + a[1].square().sum()).backward()
    (b[0].sum() + b[1].square().sum()).backward()
    pr = dict(ref.named_parameters())
    assert all(torch.equal(p.grad, pr[k].grad) for k, p in port.named_parameters())


@needs_ref
def test_multitask_loss_matches_live_reference():
    R = _ref()
    for seed in (10, 11, 12):
        loc, cls, true, u = _loss_case(seed)
        labs = []
        want_port = P.multitask_loss(loc, cls, true, (128, 128), u, labels_out=labs)
        orig = torch.multinomial

        def keyed(w, m, replacement=False, u=u):
            key = torch.where(w > 0, u[0], torch.full_like(u[0], float("inf")))
            return torch.sort(key, stable=True)[1][:m]
        torch.multinomial = keyed
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                L = R.MultiTaskLoss()
                want = L(loc, cls, true, (128, 128))
                _, labels = L.get_positive_samples_and_classification_tensor(loc, true)
        finally:
            torch.multinomial = orig
        assert torch.equal(labels, labs[0])
        assert float(want) == pytest.approx(float(want_port), rel=1e-6)


def test_dropin_module_mirrors_reference_interface():
    import inspect

    from tpgan_b200.MobileNetV2 import InvertedResidual, MobileNetV2, MultiTaskLoss, SSDHead
    torch.manual_seed(0)
    net = MobileNetV2()
    torch.manual_seed(0)
    port = P.MobileNetV2Port()
    sa, sb = net.state_dict(), port.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and len(sa) == 356
    assert all(torch.equal(sa[k], sb[k]) for k in sa)        # bit-identical seeded initialisation
    assert sum(p.numel() for p in net.parameters()) == 7_676_334
    assert MobileNetV2.num_points(128, 128) == 394
    assert list(inspect.signature(MobileNetV2.forward).parameters) == ["self", "x", "use_dropout"]
    assert list(inspect.signature(InvertedResidual.__init__).parameters) == ["self", "inp", "oup", "stride", "expand_ratio"]
    assert list(inspect.signature(SSDHead.__init__).parameters) == ["self", "num_of_out_classes"]
    assert list(inspect.signature(MultiTaskLoss.forward).parameters)[:5] == ["self", "locations_pred", "classifications_pred",
                                                                            "locations_true", "image_size"]
    L = MultiTaskLoss()
    assert (L.alpha, L.beta, L.distance_threshold_ratio, L.ratio_non_background) == (30.0, 0.1, 0.1, 5.0)  # config.py:25-27
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 128, 128))                    # no CPU fallback


def test_sgd_port_equals_torch_sgd():
    torch.manual_seed(1)
    p0 = [torch.randn(17), torch.randn(5, 3)]
    ref = [torch.nn.Parameter(t.clone()) for t in p0]
    opt = torch.optim.SGD(ref, **P.SGD)
    mine, bufs = [t.clone() for t in p0], [None, None]
    for _ in range(3):
        gs = [torch.randn_like(t) for t in p0]
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        opt.step()
        P.sgd_nesterov_step(mine, gs, bufs)
    assert all(torch.allclose(a, b.data, rtol=1e-6, atol=1e-8) for a, b in zip(mine, ref))


# ---------------------------------------------------------------------------------------------- data parallel (gloo, world 2)
def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tpgan_b200.parallel import allreduce_sum
    torch.set_num_threads(2)
    torch.manual_seed(0)
    net = P.MobileNetV2Port()

    def shard_grad(r):
        x, true, u = P.make_batch(2, seed=1234 + r)          # bench.py's per-rank seeding
        P.pretrain_step(net, x, true, u, None)
        return torch.cat([p.grad.flatten() for p in net.parameters()])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    flat = shard_grad(rank)
    allreduce_sum(flat)      # call-ordered (async_op=False): complete on return for gloo
    flat *= 1.0 / world                                      # PretrainTrainer's grad_scale of the SGD kernel
    if rank == 0:
        want = torch.zeros_like(flat)
        for r in range(world):                               # "replicas with local BatchNorm statistics" 
            net.load_state_dict(sd)
            want += shard_grad(r) / world
        ret["err"] = float((flat - want).norm() / want.norm())
    dist.destroy_process_group()


def test_gloo_world2_pretrain_gradient_is_the_mean_of_the_replicas():
    import socket

    import torch.multiprocessing as tmp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = tmp.Manager()
    ret = mgr.dict()
    tmp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["err"] < 1e-6, ret["err"]


def test_save_model_writes_a_reference_loadable_checkpoint(tmp_path):
    """save_model (UtilityMethods.py:58-76) on the drop-in module -> the file loads into the reference-layout network."""
    from tpgan_b200.MobileNetV2 import MobileNetV2
    from tpgan_b200.UtilityMethods import save_model
    torch.manual_seed(3)
    net = MobileNetV2()
    path = save_model(net, str(tmp_path / "ckpt"), 4)
    assert path.endswith("model_epoch_4.pth")
    port = P.MobileNetV2Port()
    port.load_state_dict(torch.load(path), strict=True)
    assert all(torch.equal(a, b) for a, b in zip(port.state_dict().values(), net.state_dict().values()))
    if os.path.isfile(os.path.join(REF, "MobileNetV2.py")):
        ref = _ref().MobileNetV2()
        ref.load_state_dict(torch.load(path), strict=True)


# ---------------------------------------------------------------------------------------------- decoder / accuracy (row f4)
def test_decoder_known_answer():
    """Temp.py:7-29,55-61 with nms_distance_threshold=30: one detection, class 1, confidence 0.5148, point (370, 150)."""
    loc, cls, _ = _temp_py_inputs()
    count, score, point = P.decode_sample(loc[0], cls[0], 0.5, 1, 30.0)
    assert count.tolist() == [0, 1, 0, 0, 0]
    assert float(score[1, 0]) == pytest.approx(0.5148, abs=5e-5) and point[1, 0].tolist() == [370.0, 150.0]


@needs_ref
def test_decoder_and_accuracy_match_live_reference():
    R = _ref()
    import Pretrain as PT     # has a __main__ guard; importing defines _calculate_accuracy only
    for seed in range(4):
        g = torch.Generator().manual_seed(seed)
        loc = torch.rand((1, 394, 2), generator=g) * 128
        cls = torch.randn((1, 394, 5), generator=g) * 3
        for top_k, thr, nms in ((1, 0.5, 20), (3, 0.6, 15), (8, 0.3, 40)):
            live = R.MultiTaskDecoder(thr, top_k, nms)(loc, cls)[0]
            count, score, point = P.decode_sample(loc[0], cls[0], thr, top_k, float(nms))
            mine = [(c, float(score[c, t]), point[c, t].tolist()) for c in range(5) for t in range(int(count[c]))]
            assert len(mine) == len(live)
            for (c, s, p), (lc, ls, lp) in zip(mine, live):
                assert c == lc and abs(s - float(ls)) < 1e-7 and p == lp.tolist()
    g = torch.Generator().manual_seed(3)          # a case where every class has exactly one detection (the reference's
    loc = torch.rand((1, 50, 2), generator=g) * 128    # accuracy only works then)
    cls = torch.randn((1, 50, 5), generator=g) * 6
    true = torch.rand((1, 8), generator=g) * 128
    live = R.MultiTaskDecoder(0.5, 1, 20)(loc, cls)[0]
    assert [c for c, _, _ in live] == [0, 1, 2, 3, 4]
    count, _, point = P.decode_sample(loc[0], cls[0], 0.5, 1, 20.0)
    assert P.accuracy_sample(count, point, true[0]) == pytest.approx(PT._calculate_accuracy(live, true), abs=1e-6)
