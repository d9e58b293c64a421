"""GPU parity of the bf16 (tcgen05.mma kind::f16) variants of the convolution kernels, through the C ABI, against ATen CPU
fp32 (BASELINE.json configs[2]; SURVEY.md 8c tolerance for bf16: 1e-2).

The operands handed to the kernels are made bf16-representable on the host first.  Every product of two bf16 numbers is
exact in fp32 and the tensor cores accumulate in fp32, so against the fp32 reference ON THE SAME ROUNDED OPERANDS the fp32
result of the epilogue differs by summation order only: the tolerance here is 2e-5, four hundred times sharper than the
format's own 1e-2 and sharp enough to catch a single mis-routed tap, channel chunk, K step or swizzle phase.  The bf16 twin
written by the same epilogue must equal the fp32 result rounded to bf16 (<= 2^-8 relative per element)."""
import pytest
import torch
import torch.nn.functional as F

from test_conv_gpu import ADJOINT_CASES, CONV_CASES, DECONV_CASES, MOBILENET_CASES, _mk, rel

pytestmark = pytest.mark.gpu

TOL = 2e-5
TOL16 = 4e-3      # ||bf16(v) - v|| / ||v|| <= 2^-9 * sqrt(mean) ~ 2.3e-3
GD_CASES = [c for c in CONV_CASES if c not in MOBILENET_CASES]
FWD_CASES = [c for c in GD_CASES if c[4] <= 128]


def q(t):
    """Round to bf16 (round-to-nearest-even) and back: the values the tensor cores see."""
    return t.bfloat16().float()


@pytest.fixture()
def arena():
    from tpgan_b200 import ops
    a = ops.Arena("cuda", shadow=True)
    with ops.use_arena(a):
        yield a


def _act16(t, ops):
    """NCHW host tensor -> fp32 NHWC buffer in the shadowed arena + its bf16 twin (tpgan_cast_bf16)."""
    n, c, h, w = t.shape
    a = ops.Act.empty(n, h, w, c).from_nchw(t.cuda(), round_tf32=False)
    ops.cast_bf16(a)
    return a


def _pack16(w, kind, ops):
    pk = ops.pack_weights(w.cuda(), kind, round_tf32=False)
    pk16 = ops.alloc_packed16(pk)
    ops.cast_packed(pk, pk16)
    return pk16


def _twin_nchw(a):
    return a.twin().float().permute(0, 3, 1, 2).contiguous()


def test_cast_kernels(arena):
    from tpgan_b200 import ops
    x = _mk(3, 75, 9, 11, 1)
    a = _act16(x, ops)
    torch.cuda.synchronize()
    assert torch.equal(_twin_nchw(a).cpu(), q(x))
    wide = ops.Act.empty(2, 6, 6, 208)
    part = wide.slice(140, 64)
    y = _mk(2, 64, 6, 6, 2)
    part.from_nchw(y.cuda())
    ops.cast_bf16(part)
    torch.cuda.synchronize()
    assert torch.equal(_twin_nchw(part).cpu(), q(y))
    assert float(wide.twin().float().abs().sum() - part.twin().float().abs().sum()) == 0.0   # nothing outside the slice
    w = _mk(30, 75, 3, 3, 3)
    pk = ops.pack_weights(w.cuda(), ops.CONV_FWD, round_tf32=False)
    pk16 = ops.alloc_packed16(pk)
    ops.cast_packed(pk, pk16)
    torch.cuda.synchronize()
    assert pk16.k_pad % 64 == 0 and pk16.data.dtype == torch.bfloat16
    assert torch.equal(pk16.data[:, :, :pk.k_pad].float().cpu(), q(pk.data.cpu()))
    assert float(pk16.data[:, :, pk.k_pad:].float().abs().sum()) == 0.0


@pytest.mark.parametrize("case", FWD_CASES)
def test_conv_fwd_bf16(case, arena):
    from tpgan_b200 import _lib, ops
    n, cin, cout, h, w, k, s, p = case
    x = q(_mk(n, cin, h, w, 1))
    wt = q(_mk(cout, cin, k, k, 2) * (1.0 / (cin * k * k) ** 0.5))
    b = _mk(1, cout, 1, 1, 3).flatten()
    ref = F.leaky_relu(F.conv2d(x, wt, b, stride=s, padding=p), 0.01)
    xa = _act16(x, ops)
    out = ops.Act.empty(n, ref.shape[2], ref.shape[3], cout)
    ops.conv2d(ops.CONV_FWD, xa, out, _pack16(wt, ops.CONV_FWD, ops), k, s, p, bias=b.cuda(), slope=0.01,
               epilogue=ops.EPI_LEAKY, bf16=True)
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    assert rel(out.to_nchw(), ref) < TOL, rel(out.to_nchw(), ref)
    assert rel(_twin_nchw(out), ref) < TOL16, rel(_twin_nchw(out), ref)


@pytest.mark.parametrize("case", FWD_CASES)
def test_conv_dgrad_bf16(case, arena):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    if s == 2 and (h % 2 or w % 2):
        pytest.skip("stride-2 dgrad needs even input")
    wt = q(_mk(cout, cin, k, k, 2) * (1.0 / (cout * k * k) ** 0.5))
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = q(_mk(n, cout, ho, wo, 5))
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt, dy, stride=s, padding=p)
    dx = ops.Act.empty(n, h, w, cin)
    ops.conv2d(ops.CONV_DGRAD, _act16(dy, ops), dx, _pack16(wt, ops.CONV_DGRAD, ops), k, s, p, bf16=True)
    torch.cuda.synchronize()
    assert rel(dx.to_nchw(), ref) < TOL, rel(dx.to_nchw(), ref)
    assert rel(_twin_nchw(dx), ref) < TOL16


@pytest.mark.parametrize("case", GD_CASES)
def test_conv_wgrad_bf16(case, arena):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    x = q(_mk(n, cin, h, w, 1))
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = q(_mk(n, cout, ho, wo, 5))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=p)
    dw = ops.alloc_packed(ops.CONV_FWD, (cout, cin, k, k))
    ops.wgrad(ops.CONV_FWD, _act16(x, ops), _act16(dy, ops), dw, k, s, p, bf16=True)
    got = torch.zeros((cout, cin, k, k), device="cuda")
    ops.unpack_weights(dw, got, ops.CONV_FWD)
    torch.cuda.synchronize()
    assert rel(got, ref) < TOL, rel(got, ref)


@pytest.mark.parametrize("case", DECONV_CASES)
def test_deconv_bf16(case, arena):
    """ConvTranspose2d forward, input gradient and weight gradient (phase-decomposed gather form, hole phases of k3 s4)."""
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p, op = case
    x = q(_mk(n, cin, h, w, 1)).requires_grad_(True)
    wt = q(_mk(cin, cout, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)).requires_grad_(True)
    b = _mk(1, cout, 1, 1, 3).flatten()
    y = F.conv_transpose2d(x, wt, b, stride=s, padding=p, output_padding=op)
    ref = F.relu(y)
    xa = _act16(x.detach(), ops)
    out = ops.Act.empty(n, ref.shape[2], ref.shape[3], cout)
    if k == 8:  # deconv_8 (D_and_G_model.py:218): 1x1 input -> a GEMM with N = (r, s, co); run as a 1x1 conv
        w_lin = wt.detach().permute(2, 3, 1, 0).reshape(k * k * cout, cin, 1, 1).contiguous()
        flat = ops.Act(out.buf.view(n, 1, 1, k * k * cout))
        ops.conv2d(ops.CONV_FWD, xa, flat, _pack16(w_lin, ops.CONV_FWD, ops), 1, 1, 0, bias=b.repeat(k * k).cuda(), slope=0.0,
                   epilogue=ops.EPI_LEAKY, bf16=True)
        torch.cuda.synchronize()
        assert rel(out.to_nchw(), ref) < TOL
        return
    ops.conv2d(ops.DECONV_FWD, xa, out, _pack16(wt.detach(), ops.DECONV_FWD, ops), k, s, p, bias=b.cuda(), slope=0.0,
               epilogue=ops.EPI_LEAKY, bf16=True)
    dy = q(_mk(*y.shape, 7))
    y.backward(dy)
    dya = _act16(dy, ops)
    dx = ops.Act.empty(n, h, w, cin)
    ops.conv2d(ops.DECONV_DGRAD, dya, dx, _pack16(wt.detach(), ops.DECONV_DGRAD, ops), k, s, p, bf16=True)
    dw = ops.alloc_packed(ops.DECONV_FWD, (cin, cout, k, k))
    ops.wgrad(ops.DECONV_FWD, xa, dya, dw, k, s, p, bf16=True)
    got = torch.zeros((cin, cout, k, k), device="cuda")
    ops.unpack_weights(dw, got, ops.DECONV_FWD)
    torch.cuda.synchronize()
    e0, e1, e2 = rel(out.to_nchw(), ref), rel(dx.to_nchw(), x.grad), rel(got, wt.grad)
    assert e0 < TOL and e1 < TOL and e2 < TOL, (e0, e1, e2)
    assert rel(_twin_nchw(out), ref) < TOL16


def test_epilogue_residual_mask_and_slices_bf16(arena):
    """ResidualBlock tail act(conv(h) + x) with fp32 addends, the fused activation backward with two addends and a
    per-channel mask, reading a channel slice and writing (fp32 + twin) into a slice that starts on a 4- but not
    8-channel boundary, and the twin-only output (no fp32 copy)."""
    from tpgan_b200 import ops
    n, c, h, w, k = 2, 80, 64, 64, 5
    hmid, x = q(_mk(n, c, h, w, 1)), _mk(n, c, h, w, 2)
    wt = q(_mk(c, c, k, k, 3) * (1.0 / (c * k * k) ** 0.5))
    b = _mk(1, c, 1, 1, 4).flatten()
    ref = F.leaky_relu(F.conv2d(hmid, wt, b, padding=2) + x, 0.01)
    out = ops.Act.empty(n, h, w, c)
    xa = ops.Act.empty(n, h, w, c).from_nchw(x.cuda())
    pw = _pack16(wt, ops.CONV_FWD, ops)
    ops.conv2d(ops.CONV_FWD, _act16(hmid, ops), out, pw, k, 1, 2, bias=b.cuda(), add1=xa, slope=0.01, epilogue=ops.EPI_LEAKY,
               bf16=True)
    torch.cuda.synchronize()
    assert rel(out.to_nchw(), ref) < TOL and rel(_twin_nchw(out), ref) < TOL16
    out2 = ops.Act.empty(n, h, w, c)
    ops.conv2d(ops.CONV_FWD, _act16(hmid, ops), out2, pw, k, 1, 2, bias=b.cuda(), add1=xa, slope=0.01, epilogue=ops.EPI_LEAKY,
               bf16=True, out32=False)
    torch.cuda.synchronize()
    assert float(out2.buf.abs().max()) == 0.0 and torch.equal(out2.twin(), out.twin())
    dy, a1, a2, msrc = q(_mk(n, c, h, w, 5)), _mk(n, c, h, w, 6), _mk(n, c, h, w, 7), _mk(n, c, h, w, 8)
    slopes = torch.where(torch.arange(c) % 3 == 0, 0.0, 0.01)
    slopes[5] = 1.0
    g = torch.nn.grad.conv2d_input((n, c, h, w), wt, dy, padding=2) + a1 + a2
    refm = torch.where(msrc > 0, g, g * slopes.view(1, c, 1, 1))
    dx = ops.Act.empty(n, h, w, c)
    f32 = lambda t: ops.Act.empty(*[t.shape[i] for i in (0, 2, 3, 1)]).from_nchw(t.cuda())
    ops.conv2d(ops.CONV_DGRAD, _act16(dy, ops), dx, _pack16(wt, ops.CONV_DGRAD, ops), k, 1, 2, add1=f32(a1), add2=f32(a2),
               mask=f32(msrc), slopes=slopes.cuda(), epilogue=ops.EPI_MASK, bf16=True)
    torch.cuda.synchronize()
    assert rel(dx.to_nchw(), refm) < TOL and rel(_twin_nchw(dx), refm) < TOL16
    # channel slices
    wide_in = ops.Act.empty(n, 32, 32, 160)
    xs = q(_mk(n, 128, 32, 32, 1))
    wide_in.slice(32, 128).from_nchw(xs.cuda())
    ops.cast_bf16(wide_in.slice(32, 128))
    w3 = q(_mk(64, 128, 3, 3, 2) * 0.03)
    ref3 = F.conv2d(xs, w3, None, padding=1)
    wide_out = ops.Act.empty(n, 32, 32, 208)
    ops.conv2d(ops.CONV_FWD, wide_in.slice(32, 128), wide_out.slice(140, 64), _pack16(w3, ops.CONV_FWD, ops), 3, 1, 1, bf16=True)
    torch.cuda.synchronize()
    assert rel(wide_out.slice(140, 64).to_nchw(), ref3) < TOL
    assert rel(_twin_nchw(wide_out.slice(140, 64)), ref3) < TOL16
    assert float(wide_out.slice(0, 140).to_nchw().abs().max()) == 0.0 and float(wide_out.slice(204, 4).to_nchw().abs().max()) == 0.0
    assert float(wide_out.slice(0, 140).twin().float().abs().max()) == 0.0


def test_grouped_local_pathway_shapes_bf16(arena):
    """Four problems with the local-pathway patch sizes in one launch (D_and_G_model.py:390-393), forward and weight gradient."""
    from tpgan_b200 import ops
    shapes = [(40, 40), (40, 40), (32, 40), (32, 48)]
    n, cin, cout = 3, 64, 128
    args, wargs, refs, outs, dws, keep = [], [], [], [], [], []
    for i, (h, w) in enumerate(shapes):
        x = q(_mk(n, cin, h, w, 10 + i))
        wt = q(_mk(cout, cin, 3, 3, 20 + i) * 0.04)
        dy = q(_mk(n, cout, h, w, 40 + i))
        refs.append((F.conv2d(x, wt, None, padding=1), torch.nn.grad.conv2d_weight(x, tuple(wt.shape), dy, padding=1)))
        out = ops.Act.empty(n, h, w, cout)
        pw, xa, dya = _pack16(wt, ops.CONV_FWD, ops), _act16(x, ops), _act16(dy, ops)
        dw = ops.alloc_packed(ops.CONV_FWD, tuple(wt.shape))
        keep += [xa, pw, dya]
        args.append(ops.conv_args(ops.CONV_FWD, xa, out, pw, 3, 1, 1, bf16=True))
        wargs.append(ops.wgrad_args(ops.CONV_FWD, xa, dya, dw, 3, 1, 1, bf16=True))
        outs.append(out)
        dws.append(dw)
    ops.conv2d_grouped(args)
    ops.wgrad_grouped(wargs)
    torch.cuda.synchronize()
    for o, dw, (r, rw) in zip(outs, dws, refs):
        got = torch.zeros(tuple(rw.shape), device="cuda")
        ops.unpack_weights(dw, got, ops.CONV_FWD)
        assert rel(o.to_nchw(), r) < TOL and rel(got, rw) < TOL, (rel(o.to_nchw(), r), rel(got, rw))


@pytest.mark.parametrize("case", ADJOINT_CASES)
def test_full_size_adjoint_identities_bf16(case, arena):
    """BASELINE batch size (32): <conv(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(x, dy)> with bf16-representable operands
    (exact products, fp32 accumulation): the three kernels of a layer are adjoint to fp32 summation order."""
    from tpgan_b200 import ops
    cin, cout, k, s, p, H = case
    B = 32
    Ho = (H + 2 * p - k) // s + 1
    g = torch.Generator(device="cuda").manual_seed(123)
    rnd = lambda *shape, scale=1.0: q((torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale)
    x, dy = ops.Act.empty(B, H, H, cin), ops.Act.empty(B, Ho, Ho, cout)
    x.buf[..., :cin].copy_(rnd(B, H, H, cin))
    dy.buf[..., :cout].copy_(rnd(B, Ho, Ho, cout))
    ops.cast_bf16(x)
    ops.cast_bf16(dy)
    w = rnd(cout, cin, k, k, scale=0.05)
    y, dx = ops.Act.empty(B, Ho, Ho, cout), ops.Act.empty(B, H, H, cin)
    dw = ops.alloc_packed(ops.CONV_FWD, tuple(w.shape))
    ops.conv2d(ops.CONV_FWD, x, y, _pack16(w, ops.CONV_FWD, ops), k, s, p, bf16=True)
    ops.conv2d(ops.CONV_DGRAD, dy, dx, _pack16(w, ops.CONV_DGRAD, ops), k, s, p, bf16=True)
    ops.wgrad(ops.CONV_FWD, x, dy, dw, k, s, p, accumulate=False, bf16=True)
    dwr = torch.zeros_like(w)
    ops.unpack_weights(dw, dwr, ops.CONV_FWD)
    torch.cuda.synchronize()
    dot = lambda a, b: float((a.double() * b.double()).sum())
    a1 = dot(y.buf[..., :cout], dy.buf[..., :cout])
    a2 = dot(x.buf[..., :cin], dx.buf[..., :cin])
    a3 = dot(w, dwr)
    bound = float(y.buf[..., :cout].double().norm() * dy.buf[..., :cout].double().norm())
    assert abs(a1 - a2) <= 1e-4 * bound and abs(a1 - a3) <= 1e-4 * bound, (a1, a2, a3, bound)
    assert abs(a1) > 0 and bound > 0
