"""CPU tests that PIN THE ORACLE: the fp32 port (oracle/model_port.py) and the oracle step against
  (a) the live reference modules (only where /root/reference exists - the build container), and
  (b) the committed golden vectors generated from the live reference by tools/make_golden.py (everywhere).
The reference has no tests and no golden vectors of its own for the G/D path."""
import os

import numpy as np
import pytest
import torch

from oracle import model_port as mp, reference, step as ostep

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.pt")
NAMES = ("img", "left_eye", "right_eye", "nose", "mouth", "z")
needs_ref = pytest.mark.skipif(not reference.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


@pytest.fixture(scope="module")
def seeded_state():
    """Seeded weights built by OUR drop-in constructors (bit-identical to the reference's, see test below)."""
    from tpgan_b200 import D_and_G_model as M, config
    torch.manual_seed(0)
    G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"])
    D = M.Discriminator(config.D["use_batchnorm"])
    return G.state_dict(), D.state_dict()


@needs_ref
def test_reference_invariants_and_dropin_state_dict(seeded_state):
    G, D = reference.build_models(0)
    reference.self_check(G, D)
    sg, sd = seeded_state
    rg, rd = G.state_dict(), D.state_dict()
    assert list(sg.keys()) == list(rg.keys()) and list(sd.keys()) == list(rd.keys())
    for k in rg:
        assert torch.equal(sg[k], rg[k]), k       # same shapes, same RNG consumption order, same values
    for k in rd:
        assert torch.equal(sd[k], rd[k]), k


@needs_ref
def test_config_equals_reference():
    import tpgan_b200.config as c
    rc = reference.load().config
    for name in ("optimizer_param", "general", "train", "G", "D", "loss", "feature_extract_model"):
        assert getattr(c, name) == getattr(rc, name), name
    assert ostep.LOSS_W == rc.loss and ostep.LEARNING_RATE == rc.train["learning_rate"]


@needs_ref
def test_port_is_bit_exact_with_live_reference():
    G, D = reference.build_models(0)
    b = ostep.make_batch(1)
    with torch.no_grad():
        ref = G(*[b[k] for k in NAMES], False)
        got = mp.generator(G.state_dict(), *[b[k] for k in NAMES])
        for r, g in zip(ref, got):
            assert torch.equal(r, g)
        assert torch.equal(D(b["img"]), mp.discriminator(D.state_dict(), b["img"]))


def test_port_matches_golden_forward(gold, seeded_state):
    sg, sd = seeded_state
    b = ostep.make_batch(1)
    with torch.no_grad():
        got = mp.generator(sg, *[b[k] for k in NAMES])
        for r, g in zip(gold["g_forward_b1"], got):
            assert torch.allclose(r, g, rtol=0, atol=2e-6), float((r - g).abs().max())
        assert torch.allclose(gold["d_forward_b1"], mp.discriminator(sd, b["img"]), rtol=0, atol=2e-6)
        fused = mp.local_fuser([b[k] for k in NAMES[1:5]])
        assert torch.equal(fused, gold["fuser_b1"])


def test_fuser_geometry_and_clamp():
    """LocalFuser rectangles measured on the reference and the max(0, .) clamp of the zero padding."""
    assert mp.FUSE_RECTS == ((18, 19, 40, 40), (65, 18, 40, 40), (43, 47, 40, 32), (40, 72, 48, 32))
    parts = [-torch.ones(1, 2, h, w) for (l, t, w, h) in mp.FUSE_RECTS]
    assert float(mp.local_fuser(parts).abs().max()) == 0.0          # all-negative patches fuse to exactly 0
    parts = [torch.full((1, 1, h, w), float(i + 1)) for i, (l, t, w, h) in enumerate(mp.FUSE_RECTS)]
    val, idx = mp.local_fuser(parts, return_index=True)
    assert int((val > 0).sum()) == 5358                               # covered pixels (probe)
    assert float(val[0, 0, 50, 50]) == 3.0 and int(idx[0, 0, 50, 50]) == 2   # left eye / nose overlap: max wins


def test_crop_boxes_match_reference_process(gold):
    """process() (DataAndDataset.py:10-56) recorded from the live reference: boxes + PIL zero fill."""
    lms = gold["process_landmarks"].numpy()
    img = gold["process_image_u8"]
    boxes = ostep.crop_boxes(lms)
    t = img.permute(2, 0, 1)[None].float()
    ours = ostep.crop_patches(t.repeat(len(lms), 1, 1, 1), lms, fill=0.0)
    for n in range(len(lms)):
        for i, key in enumerate(("left_eye", "right_eye", "nose", "mouth")):
            ref = gold["process_crops"][n][key].permute(2, 0, 1).float()
            assert torch.equal(ours[i][n], ref), (n, key, boxes[n, i])
    # canonical left-eye landmark (39.48, 40.28) -> img[21:61, 20:60]   
    assert boxes[0, 0].tolist() == [20, 21, 60, 61]


def test_oracle_step_matches_golden(gold, seeded_state):
    """Oracle step on the port == oracle step on the live reference modules (losses and gradient fingerprints)."""
    sg, sd = seeded_state
    pg = {k: v.clone().requires_grad_(True) for k, v in sg.items()}
    pd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    Gc, Dc = ostep.port_callables(pg, pd)
    b = ostep.make_batch(2)
    g_out = Gc(b)
    ld, md = ostep.d_loss(Dc, g_out[0].detach(), b)
    gd = torch.autograd.grad(ld, list(pd.values()))
    lg, mg = ostep.g_loss(g_out, Dc(g_out[0]), b)
    gg = torch.autograd.grad(lg, list(pg.values()))
    m = {k: float(v) for k, v in {**md, **mg}.items()}
    for k, v in gold["step_b2_metrics"].items():
        assert abs(m[k] - v) <= 1e-4 * abs(v) + 1e-6, (k, m[k], v)
    for (n, _), g in zip(pg.items(), gg):
        ref = gold["step_b2_g_gradnorm"][n]
        assert abs(float(g.norm()) - ref) <= 2e-3 * ref + 1e-7, (n, float(g.norm()), ref)
        if n in gold["step_b2_g_grads"]:
            r = gold["step_b2_g_grads"][n]
            assert float((g - r).norm() / r.norm()) < 2e-3, n
    for (n, _), g in zip(pd.items(), gd):
        ref = gold["step_b2_d_gradnorm"][n]
        assert abs(float(g.norm()) - ref) <= 2e-3 * ref + 1e-7, (n, float(g.norm()), ref)


def test_tf32_rounding_emulation():
    x = torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -11 - 2 ** -20, -1.0 - 2 ** -11, 3.14159, 0.0, float("inf")])
    r = mp.tf32_rna(x)
    assert r.tolist()[:5] == [1.0 + 2 ** -10, 1.0, -1.0 - 2 ** -10, 3.140625, 0.0] and torch.isinf(r[5])


def test_flops_table_matches_survey():
    """Algorithmic forward FLOPs per image from the layer shapes: 176.56 (G) / 1.298 (D) GFLOP."""
    from oracle import layer_table
    f = layer_table.flops()
    assert abs(f["G"] / 1e9 - 176.56) < 0.02 and abs(f["D"] / 1e9 - 1.298) < 0.002, f


def test_identity_port_is_the_canonical_resnet18():
    """oracle/identity_port.py restates the ResNet18 the reference's ResNet.py describes but cannot build.
    Its wiring is pinned against torchvision's resnet18 (the canonical network the reference file follows: 7x7/2 stem,
    MaxPool(3,2,1), [2,2,2,2] basic blocks, stage strides 1,2,2,2, 1x1 projection shortcuts, global average pool, FC):
    same weights -> same logits on a 128x128 input."""
    tv = pytest.importorskip("torchvision")
    from oracle import identity_port as ip
    torch.manual_seed(0)
    ref = tv.models.resnet18(num_classes=37).eval()
    g = torch.Generator().manual_seed(1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    src = ref.state_dict()
    sd = {}

    def bn(dst, s):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[f"{dst}.{k}"] = src[f"{s}.{k}"]
    sd["conv1.0.weight"] = src["conv1.weight"]
    bn("conv1.1", "bn1")
    for s in range(4):
        for b in range(2):
            a, d = f"layer{s + 1}.{b}", f"sections.{s}.{b}"
            sd[d + ".conv_a.0.weight"] = src[a + ".conv1.weight"]
            bn(d + ".conv_a.1", a + ".bn1")
            sd[d + ".conv_b.0.weight"] = src[a + ".conv2.weight"]
            bn(d + ".conv_b.1", a + ".bn2")
            if a + ".downsample.0.weight" in src:
                sd[d + ".shortcut.0.weight"] = src[a + ".downsample.0.weight"]
                bn(d + ".shortcut.1", a + ".downsample.1")
    sd["FC.0.weight"], sd["FC.0.bias"] = src["fc.weight"], src["fc.bias"]
    x = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        want = ref(x)
        got, fc0, pooled = ip.resnet18_128(sd, x)
    assert fc0 is None and pooled.shape == (2, 512)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())


def test_identity_modules_mirror_reference_interface():
    """tpgan_b200.ResNet / FeatureExtract keep the reference's constructor and forward signatures (ResNet.py:6-7,80;
    FeatureExtract.py:6,40) and the state_dict key structure of its layer factories."""
    from tpgan_b200.FeatureExtract import FeatureExtractModel
    from tpgan_b200.ResNet import BasicBlock, ResNet18
    net = FeatureExtractModel("resnet", 347, residualBlock=BasicBlock, feature_layer_dim_before_FC=256)
    keys = set(net.state_dict())
    assert "base_model.conv1.0.weight" in keys and "base_model.conv1.1.running_mean" in keys
    assert "base_model.FC0.0.weight" in keys and "base_model.FC.0.weight" in keys
    assert net.base_model.FC[0].out_features == 347 and net.base_model.FC[0].in_features == 256
    assert sum(p.numel() for p in ResNet18(BasicBlock, 1000).parameters()) == 11689512   # canonical ResNet18
    with pytest.raises(ValueError):
        FeatureExtractModel("vgg")
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 3, 128, 128))   # CPU tensor: the product path has no CPU fallback


@needs_ref
def test_batchnorm_step_golden_is_the_live_reference():
    """tests/golden/step_bn_golden.pt (the fused BatchNorm-generator step's fixture, tools/make_golden_bn_step.py) regenerated
    in process from the live reference: Generator(use_batchnorm=True) + BatchNorm-free Discriminator, one oracle step."""
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "step_bn_golden.pt"), weights_only=False)
    ns = reference.load()
    torch.manual_seed(0)
    G, D = ns.DG.Generator(64, 347, True, False), ns.DG.Discriminator(False)
    G.train()
    b = ostep.make_batch(2, seed=3)
    m = ostep.train_step(lambda bb: G(bb["img"], bb["left_eye"], bb["right_eye"], bb["nose"], bb["mouth"], bb["z"], False), D,
                         list(G.parameters()), list(D.parameters()), torch.optim.Adam(G.parameters(), lr=1e-4),
                         torch.optim.Adam(D.parameters(), lr=1e-4), b, step_optim=False)
    for k, v in gold["metrics"].items():
        assert abs(m[k] - v) <= 1e-5 * abs(v) + 1e-6, (k, m[k], v)
    sd = G.state_dict()
    for k, v in gold["running"].items():
        assert torch.allclose(sd[k], v, rtol=1e-5, atol=1e-7), k
    for k, g in gold["bn_grads"].items():
        assert torch.allclose(dict(G.named_parameters())[k].grad, g, rtol=1e-4, atol=1e-6), k
