"""GPU parity of the HBM-bound kernels (through the C ABI) against plain PyTorch on the CPU / the oracle.
Index work (crop boxes, stitch argmax) is bit-exact; fp32 elementwise work is exact or within a few ulp."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.pt")


def _act(t):
    from tpgan_b200 import ops
    n, c, h, w = t.shape
    return ops.Act.empty(n, h, w, c).from_nchw(t.cuda())


def _rand(*s, seed=0):
    return torch.rand(*s, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def test_patch_crop_bit_exact_incl_out_of_image():
    from oracle import step as ostep
    from tpgan_b200 import ops
    gold = torch.load(GOLD, weights_only=False)
    lms = gold["process_landmarks"]                       # incl. a landmark set whose boxes leave the image
    n = lms.shape[0]
    img = (gold["process_image_u8"].permute(2, 0, 1)[None].float() / 255 * 2 - 1).repeat(n, 1, 1, 1)
    ia = _act(img)
    outs = [ops.Act.empty(n, h, w, 3) for (w, h) in ostep.PATCH_WH]
    boxes = torch.zeros((n, 4, 4), dtype=torch.int32, device="cuda")
    ops.patch_crop(ia, lms.cuda().contiguous(), outs, boxes, fill=-1.0)
    torch.cuda.synchronize()
    assert (boxes.cpu().numpy() == ostep.crop_boxes(lms.numpy())).all()
    ref = ostep.crop_patches(img, lms.numpy(), fill=-1.0)
    for o, r in zip(outs, ref):
        assert torch.equal(o.to_nchw().cpu(), r)
    # and against the reference's own PIL crops (uint8 domain): -1 fill <-> PIL's 0
    for k in range(n):
        for i, key in enumerate(("left_eye", "right_eye", "nose", "mouth")):
            pil = gold["process_crops"][k][key].permute(2, 0, 1).float() / 255 * 2 - 1
            assert torch.allclose(outs[i].to_nchw().cpu()[k], pil, atol=1e-6)


@pytest.mark.parametrize("c", [3, 64])
def test_local_fuse_forward_backward_bit_exact(c):
    from oracle import model_port as mp
    from tpgan_b200 import ops
    n = 3
    parts = [_rand(n, c, h, w, seed=i) for i, (l, t, w, h) in enumerate(mp.FUSE_RECTS)]
    parts[0][:, :, :5] = 0.0                               # exact ties with the zero padding
    parts[2][:, :, 0:8, 0:8] = parts[0][:, :, 28:36, 25:33]  # exact ties between two patches in their overlap
    val, idx = mp.local_fuser(parts, return_index=True)
    pa = [_act(p) for p in parts]
    out = ops.Act.empty(n, 128, 128, c)
    am = torch.empty((n, 128, 128, c), dtype=torch.uint8, device="cuda")
    ops.local_fuse(pa, out, am)
    torch.cuda.synchronize()
    assert torch.equal(out.to_nchw().cpu(), val)
    assert torch.equal(am.cpu().permute(0, 3, 1, 2).long(), idx)       # first index wins ties, like torch.max
    # backward: gradient routed to the arg-max source
    pr = [p.clone().requires_grad_(True) for p in parts]
    g = _rand(n, c, 128, 128, seed=9)
    mp.local_fuser(pr).backward(g)
    dp = [ops.Act.empty(n, p.shape[2], p.shape[3], c) for p in parts]
    ops.local_fuse_backward(_act(g), am, dp, accumulate=False)
    torch.cuda.synchronize()
    for d, p in zip(dp, pr):
        assert torch.equal(d.to_nchw().cpu(), p.grad)


def test_image_losses_values_and_gradient():
    from oracle import step as ostep
    from tpgan_b200 import ops
    B = 3
    fake = _rand(B, 3, 128, 128, seed=1).requires_grad_(True)
    gt = _rand(B, 3, 128, 128, seed=2)
    b = dict(img_frontal=gt, img64_frontal=F.avg_pool2d(gt, 2), img32_frontal=F.avg_pool2d(gt, 4))
    w = ostep.LOSS_W
    pixel, sym, tv = ostep.image_terms(fake, b)
    total = w["weight_pixelwise"] * pixel + w["weight_symmetry"] * sym + w["weight_total_varation"] * tv
    total.backward()
    n128, n64, n32 = B * 3 * 128 * 128, B * 3 * 64 * 64, B * 3 * 32 * 32
    wp, ws, wt = w["weight_pixelwise"], w["weight_symmetry"], w["weight_total_varation"]
    coeffs = [wp * w["weight_128"] / n128, wp * w["weight_64"] / n64, wp * w["weight_32"] / n32,
              ws * w["weight_128"] / n128, ws * w["weight_64"] / n64, ws * w["weight_32"] / n32,
              wt / (B * 3 * 127 * 128), wt / (B * 3 * 128 * 127)]
    fa = _act(fake.detach())
    df = ops.Act.empty(B, 128, 128, 3)
    sums = torch.zeros(8, device="cuda")
    ops.image_losses(fa, _act(gt), _act(b["img64_frontal"]), _act(b["img32_frontal"]), df, coeffs, sums)
    torch.cuda.synchronize()
    s = sums.cpu().tolist()
    got_pixel = w["weight_128"] * s[0] / n128 + w["weight_64"] * s[1] / n64 + w["weight_32"] * s[2] / n32
    got_sym = w["weight_128"] * s[3] / n128 + w["weight_64"] * s[4] / n64 + w["weight_32"] * s[5] / n32
    got_tv = s[6] / (B * 3 * 127 * 128) + s[7] / (B * 3 * 128 * 127)
    assert abs(got_pixel - float(pixel)) < 1e-5 * float(pixel)
    assert abs(got_sym - float(sym)) < 1e-5 * float(sym)
    assert abs(got_tv - float(tv)) < 1e-5 * float(tv)
    g, r = df.to_nchw().cpu(), fake.grad
    # the gradient is a sum of signs: identical except where an L1 argument is within rounding of zero
    bad = ((g - r).abs() > 1e-9).float().mean()
    assert float(bad) < 1e-3 and float((g - r).norm() / r.norm()) < 2e-2


def test_image_losses_tiled_matches_generic(monkeypatch):
    """The shared-memory tiled kernel (the one the step's float4 layout takes) against the per-element kernel: gradients
    bit-identical, sums to fp32 addition order."""
    from tpgan_b200 import ops
    B = 5
    fake, gt = _rand(B, 3, 128, 128, seed=11), _rand(B, 3, 128, 128, seed=12)
    t64, t32 = F.avg_pool2d(gt, 2), F.avg_pool2d(gt, 4)
    coeffs = [1e-6 * (i + 1) for i in range(8)]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("TPGAN_IMAGE_LOSSES_GENERIC", mode)
        df = ops.Act.empty(B, 128, 128, 3)
        sums = torch.zeros(8, device="cuda")
        ops.image_losses(_act(fake), _act(gt), _act(t64), _act(t32), df, coeffs, sums)
        torch.cuda.synchronize()
        res[mode] = (df.to_nchw().cpu(), sums.cpu())
    assert torch.equal(res["0"][0], res["1"][0])
    assert torch.allclose(res["0"][1], res["1"][1], rtol=1e-5)


def test_l1_ce_maxout_lerp_mul():
    from tpgan_b200 import ops
    a, b = _rand(4, 3, 32, 48, seed=1), _rand(4, 3, 32, 48, seed=2)
    ar = a.clone().requires_grad_(True)
    (3.0 * (ar - b).abs().mean()).backward()
    da, tot = ops.Act.empty(4, 32, 48, 3), torch.zeros(1, device="cuda")
    ops.l1_loss(_act(a), _act(b), da, 3.0 / a.numel(), tot)
    assert abs(float(tot.cpu()) / a.numel() - float((a - b).abs().mean())) < 1e-6
    assert torch.allclose(da.to_nchw().cpu(), ar.grad, atol=1e-9)
    # cross entropy
    lg = _rand(5, 347, seed=3) * 4
    lab = torch.randint(0, 347, (5,), generator=torch.Generator().manual_seed(4))
    lr = lg.clone().requires_grad_(True)
    ce = F.cross_entropy(lr, lab)
    (10.0 * ce).backward()
    la = ops.Act.empty(5, 1, 1, 347).from_nchw(lg.view(5, 347, 1, 1).cuda())
    dl, s = ops.Act.empty(5, 1, 1, 347), torch.zeros(1, device="cuda")
    ops.softmax_ce(la, lab.cuda(), dl, 10.0 / 5, s)
    assert abs(float(s.cpu()) / 5 - float(ce)) < 1e-5
    assert torch.allclose(dl.to_nchw().cpu().view(5, 347), lr.grad, atol=1e-6)
    # maxout (MaxPool1d(2,2)) with a tie
    x = _rand(6, 512, seed=5)
    x[0, 0] = x[0, 1]
    xr = x.clone().requires_grad_(True)
    y = F.max_pool1d(xr.view(6, -1, 2), 2, 2).view(6, -1)
    gy = _rand(6, 256, seed=6)
    y.backward(gy)
    xc, yc, dx = x.cuda(), torch.empty(6, 256, device="cuda"), torch.empty(6, 512, device="cuda")
    ops.maxout2(xc, yc)
    ops.maxout2_backward(xc, gy.cuda(), dx)
    assert torch.equal(yc.cpu(), y.detach()) and torch.equal(dx.cpu(), xr.grad)
    # lerp / mul / split
    al = torch.rand(4, generator=torch.Generator().manual_seed(7))
    o = ops.Act.empty(4, 32, 48, 3)
    ops.lerp(_act(a), _act(b), al.cuda(), o)
    assert torch.allclose(o.to_nchw().cpu(), al.view(-1, 1, 1, 1) * a + (1 - al.view(-1, 1, 1, 1)) * b, atol=1e-6)
    ops.mul(_act(a), _act(b), o)
    assert torch.equal(o.to_nchw().cpu(), a * b)
    hi, lo = ops.Act.empty(4, 32, 48, 3), ops.Act.empty(4, 32, 48, 3)
    ops.split_tf32(_act(a), hi, lo)
    from oracle.model_port import tf32_rna
    assert torch.equal(hi.to_nchw().cpu(), tf32_rna(a)) and torch.equal(lo.to_nchw().cpu(), a - tf32_rna(a))


def test_reflect_pad_bias_grad_act_backward_view_copy():
    from tpgan_b200 import ops
    x = _rand(2, 12, 8, 8, seed=1)
    xr = x.clone().requires_grad_(True)
    y = F.pad(xr, (1, 0, 1, 0), mode="reflect")
    g = _rand(2, 12, 9, 9, seed=2)
    y.backward(g)
    out, dx = ops.Act.empty(2, 9, 9, 12), ops.Act.empty(2, 8, 8, 12)
    ops.reflect_pad(_act(x), out, 1, 1)
    ops.reflect_pad_backward(_act(g), dx, 1, 1, accumulate=False)
    assert torch.equal(out.to_nchw().cpu(), y.detach()) and torch.allclose(dx.to_nchw().cpu(), xr.grad, atol=1e-6)
    db = torch.zeros(12, device="cuda")
    ops.bias_grad(_act(g), db, accumulate=True)
    assert torch.allclose(db.cpu(), g.sum((0, 2, 3)), atol=1e-4)
    m = _rand(2, 12, 9, 9, seed=3)
    o = ops.Act.empty(2, 9, 9, 12)
    ops.act_backward(_act(g), _act(m), o, slope=0.01)
    assert torch.equal(o.to_nchw().cpu(), torch.where(m > 0, g, g * 0.01))
    wide = ops.Act.empty(2, 9, 9, 40)
    ops.view_copy(_act(g), wide.slice(8, 12))
    ops.view_copy(_act(g), wide.slice(8, 12), accumulate=True)
    assert torch.equal(wide.slice(8, 12).to_nchw().cpu(), 2 * g) and float(wide.slice(0, 8).to_nchw().abs().max()) == 0


def test_adam_matches_torch():
    from tpgan_b200 import ops
    p0, g = _rand(10007, seed=1), _rand(10007, seed=2) * 0.1
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1e-4)
    pc, m, v = p0.clone().cuda(), torch.zeros(10007, device="cuda"), torch.zeros(10007, device="cuda")
    for step in range(1, 4):
        p.grad = g * step
        opt.step()
        ops.adam_step(pc, (g * step).cuda(), m, v, 1e-4, 0.9, 0.999, 1e-8, 0.0, step)
    assert torch.allclose(pc.cpu(), p.detach(), atol=2e-7)


def test_gp_coeff_and_sample_ops():
    from tpgan_b200 import ops
    g = _rand(5, 3, 128, 128, seed=1) * 0.01
    ga = _act(g)
    sq, co, s = torch.zeros(5, device="cuda"), torch.zeros(5, device="cuda"), torch.zeros(1, device="cuda")
    ops.sample_sqnorm(ga, sq)
    ops.gp_coeff(sq, co, 10.0 * 2 / 5, s)
    nr = g.flatten(1).norm(dim=1)
    assert torch.allclose(sq.cpu().sqrt(), nr, rtol=1e-5)
    assert torch.allclose(co.cpu(), 4.0 * (nr - 1) / nr, rtol=1e-4)
    assert abs(float(s.cpu()) - float(((nr - 1) ** 2).sum())) < 1e-4
    u = ops.Act.empty(5, 128, 128, 3)
    ops.sample_scale(ga, co, u)
    assert torch.allclose(u.to_nchw().cpu(), co.cpu().view(-1, 1, 1, 1) * g, rtol=1e-6, atol=1e-9)
