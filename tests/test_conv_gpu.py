"""GPU parity of the tcgen05 convolution kernels (through the C ABI) against ATen CPU fp32, which is the arithmetic the
reference's conv()/deconv() factories lower to (ModificationLayer.py:101,189).

Tolerance: TF32 inputs (10-bit mantissa, round-to-nearest) with fp32 accumulation -> norm-wise relative error
||a-b||_2/||b||_2 <= 1e-3 per op (north-star tolerance for tf32)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _mk(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, c, h, w), generator=g) * 2 - 1


def _act(t, ops):
    n, c, h, w = t.shape
    return ops.Act.empty(n, h, w, c).from_nchw(t.cuda(), round_tf32=True)


CONV_CASES = [
    # n, cin, cout, h, w, k, stride, pad
    (2, 32, 16, 8, 8, 1, 1, 0),
    (2, 64, 64, 16, 16, 3, 1, 1),
    (1, 8, 64, 128, 128, 3, 1, 1),
    (2, 206, 206, 32, 32, 5, 1, 2),
    (1, 3, 64, 128, 128, 7, 1, 3),
    (2, 75, 75, 64, 64, 7, 1, 3),
    (3, 64, 128, 32, 32, 3, 2, 1),
    (2, 64, 64, 128, 128, 5, 2, 2),
    (5, 3, 64, 40, 40, 3, 1, 1),
    (5, 64, 128, 40, 40, 3, 2, 1),
    (7, 256, 512, 10, 10, 3, 2, 1),
    (7, 512, 512, 5, 5, 3, 1, 1),
    (3, 128, 64, 32, 48, 3, 1, 1),
    (4, 576, 576, 9, 9, 2, 1, 0),
    (2, 416, 416, 32, 32, 3, 1, 1),
    (4, 512, 1, 4, 4, 3, 1, 1),
    (4, 64, 3, 40, 40, 1, 1, 0),
    # full-width (W = 128) stride-1 layers: served by the row-tile kernel (rowconv.cu); H need not divide the row tile
    (2, 206, 206, 12, 128, 5, 1, 2),
    (2, 75, 75, 7, 128, 7, 1, 3),
    (3, 64, 64, 9, 128, 7, 1, 3),
    (2, 206, 64, 10, 128, 5, 1, 2),
    (2, 64, 32, 6, 128, 3, 1, 1),
    (1, 32, 3, 5, 128, 3, 1, 1),
    # slab-mode weight gradients: partial last pixel box (W = 200), even kernel, wide-P / narrow-Q (operand swap)
    (1, 64, 64, 4, 200, 3, 1, 1),
    (2, 40, 24, 6, 36, 4, 1, 2),
    (2, 160, 32, 5, 64, 5, 1, 2),
    # CTA-pair launches (tapgemm over clusters of 2) with an ODD number of 128-pixel tiles: the last pair's second CTA runs a
    # tile past the end (zero-filled loads, no stores)
    (5, 64, 64, 8, 8, 3, 1, 1),
    (3, 32, 48, 24, 24, 3, 1, 1),
    (3, 96, 208, 24, 24, 3, 2, 1),
]

# every distinct dense-conv shape of MobileNetV2-SSD at 128x128 (Pretrain path, MobileNetV2.py:28-44,103-171): the stem,
# the 1x1 expand / project convs of the 17 bottlenecks, conv2, the extra layers (3x3 stride 2 down to 1x1 maps - a 3x3
# stride-2 pad-1 conv on a 1x1 map is run as stride 1, see tpgan_b200/MobileNetV2.py::_layer_at) and the 12 SSD heads
MOBILENET_CASES = [
    (4, 3, 32, 128, 128, 3, 2, 1),
    (4, 32, 32, 64, 64, 1, 1, 0), (4, 32, 16, 64, 64, 1, 1, 0), (4, 16, 96, 64, 64, 1, 1, 0), (4, 96, 24, 32, 32, 1, 1, 0),
    (4, 24, 144, 32, 32, 1, 1, 0), (4, 144, 24, 32, 32, 1, 1, 0), (4, 144, 32, 16, 16, 1, 1, 0), (4, 32, 192, 16, 16, 1, 1, 0),
    (4, 192, 32, 16, 16, 1, 1, 0), (4, 192, 64, 8, 8, 1, 1, 0), (4, 64, 384, 8, 8, 1, 1, 0), (4, 384, 64, 8, 8, 1, 1, 0),
    (4, 384, 96, 8, 8, 1, 1, 0), (4, 96, 576, 8, 8, 1, 1, 0), (4, 576, 96, 8, 8, 1, 1, 0), (4, 576, 160, 4, 4, 1, 1, 0),
    (4, 160, 960, 4, 4, 1, 1, 0), (4, 960, 160, 4, 4, 1, 1, 0), (4, 960, 320, 4, 4, 1, 1, 0), (4, 320, 1280, 4, 4, 1, 1, 0),
    (4, 1280, 512, 4, 4, 1, 1, 0), (4, 512, 512, 4, 4, 3, 2, 1), (4, 512, 256, 2, 2, 1, 1, 0), (4, 256, 256, 2, 2, 3, 2, 1),
    (4, 256, 256, 1, 1, 3, 1, 1), (4, 256, 128, 1, 1, 1, 1, 0), (4, 128, 128, 1, 1, 3, 1, 1),
    (4, 96, 8, 8, 8, 3, 1, 1), (4, 96, 20, 8, 8, 3, 1, 1), (4, 1280, 12, 4, 4, 3, 1, 1), (4, 1280, 30, 4, 4, 3, 1, 1),
    (4, 512, 12, 2, 2, 3, 1, 1), (4, 512, 30, 2, 2, 3, 1, 1), (4, 256, 12, 1, 1, 3, 1, 1), (4, 256, 30, 1, 1, 3, 1, 1),
    (4, 128, 12, 1, 1, 3, 1, 1), (4, 128, 30, 1, 1, 3, 1, 1),
]
CONV_CASES = CONV_CASES + MOBILENET_CASES


FWD_CASES = [c for c in CONV_CASES if c[4] <= 128]   # the forward/dgrad kernels tile widths up to 128


@pytest.mark.parametrize("case", FWD_CASES)
def test_conv_fwd(case):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    x = _mk(n, cin, h, w, 1)
    wt = _mk(cout, cin, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)
    b = _mk(1, cout, 1, 1, 3).flatten()
    ref = F.leaky_relu(F.conv2d(x, wt, b, stride=s, padding=p), 0.01)
    xa = _act(x, ops)
    out = ops.Act.empty(n, ref.shape[2], ref.shape[3], cout)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
    ops.conv2d(ops.CONV_FWD, xa, out, pw, k, s, p, bias=b.cuda(), slope=0.01, epilogue=ops.EPI_LEAKY, round_tf32=True)
    torch.cuda.synchronize()
    from tpgan_b200 import _lib
    assert _lib.kernel_status() == 0
    err = rel(out.to_nchw(), ref)
    assert err < TOL, err


def test_conv_runs_over_cta_pairs():
    """Launches with two or more 128-pixel tiles run tapgemm over CTA pairs (cta_group::2) unless TPGAN_PAIR=0; a single-tile
    launch and the split-K Linear layers stay on single CTAs.  Results are the same arithmetic either way (test_conv_fwd)."""
    import os
    from tpgan_b200 import _lib, ops
    def run(n, cin, cout, h, k, p):
        x = _mk(n, cin, h, h, 1)
        wt = _mk(cout, cin, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)
        ref = F.conv2d(x, wt, None, padding=p)
        out = ops.Act.empty(n, ref.shape[2], ref.shape[3], cout)
        ops.conv2d(ops.CONV_FWD, _act(x, ops), out, ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True), k, 1, p,
                   round_tf32=True)
        torch.cuda.synchronize()
        assert _lib.kernel_status() == 0
        assert rel(out.to_nchw(), ref) < TOL
        return _lib.last_conv_kernel(), _lib.last_conv_pair()
    want = os.environ.get("TPGAN_PAIR", "1") != "0"
    assert run(5, 64, 64, 8, 3, 1) == ("tapgemm", want)        # 3 tiles: one full pair + one half-empty pair
    assert run(2, 64, 64, 8, 3, 1) == ("tapgemm", False)       # one tile


@pytest.mark.parametrize("case", FWD_CASES)
def test_conv_dgrad(case):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    if s == 2 and (h % 2 or w % 2):
        pytest.skip("stride-2 dgrad needs even input")
    wt = _mk(cout, cin, k, k, 2) * (1.0 / (cout * k * k) ** 0.5)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = _mk(n, cout, ho, wo, 5)
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt, dy, stride=s, padding=p)
    dya = _act(dy, ops)
    dx = ops.Act.empty(n, h, w, cin)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_DGRAD, round_tf32=True)
    ops.conv2d(ops.CONV_DGRAD, dya, dx, pw, k, s, p, round_tf32=True)
    torch.cuda.synchronize()
    err = rel(dx.to_nchw(), ref)
    assert err < TOL, err


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_wgrad(case):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    x = _mk(n, cin, h, w, 1)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = _mk(n, cout, ho, wo, 5)
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=p)
    xa, dya = _act(x, ops), _act(dy, ops)
    dw = ops.alloc_packed(ops.CONV_FWD, (cout, cin, k, k))
    ops.wgrad(ops.CONV_FWD, xa, dya, dw, k, s, p)
    got = torch.zeros((cout, cin, k, k), device="cuda")
    ops.unpack_weights(dw, got, ops.CONV_FWD)
    torch.cuda.synchronize()
    err = rel(got, ref)
    assert err < TOL, err


DECONV_CASES = [
    # n, cin, cout, h, w, k, stride, pad, out_pad
    (2, 512, 256, 5, 5, 3, 2, 1, 1),
    (2, 128, 64, 20, 20, 3, 2, 1, 1),
    (2, 208, 64, 64, 64, 3, 2, 1, 1),
    (3, 64, 32, 8, 8, 3, 4, 0, 1),
    (4, 320, 64, 1, 1, 8, 1, 0, 0),
    (2, 16, 8, 64, 64, 3, 2, 1, 1),
    (2, 256, 128, 8, 12, 3, 2, 1, 1),
    (2, 32, 16, 8, 40, 3, 1, 1, 0),
]


def _deconv_as_args(case):
    n, cin, cout, h, w, k, s, p, op = case
    return n, cin, cout, h, w, k, s, p, op


@pytest.mark.parametrize("case", DECONV_CASES)
def test_deconv_fwd(case):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p, op = case
    x = _mk(n, cin, h, w, 1)
    wt = _mk(cin, cout, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)
    b = _mk(1, cout, 1, 1, 3).flatten()
    ref = F.relu(F.conv_transpose2d(x, wt, b, stride=s, padding=p, output_padding=op))
    xa = _act(x, ops)
    out = ops.Act.empty(n, ref.shape[2], ref.shape[3], cout)
    if k == 8:  # deconv_8 (D_and_G_model.py:218): 1x1 input -> a GEMM with N = (r, s, co); run as a 1x1 conv
        w_lin = wt.permute(2, 3, 1, 0).reshape(k * k * cout, cin, 1, 1).contiguous()
        b_lin = b.repeat(k * k)
        pw = ops.pack_weights(w_lin.cuda(), ops.CONV_FWD, round_tf32=True)
        flat = ops.Act(out.buf.view(n, 1, 1, k * k * cout))
        ops.conv2d(ops.CONV_FWD, xa, flat, pw, 1, 1, 0, bias=b_lin.cuda(), slope=0.0, epilogue=ops.EPI_LEAKY, round_tf32=True)
    else:
        pw = ops.pack_weights(wt.cuda(), ops.DECONV_FWD, round_tf32=True)
        ops.conv2d(ops.DECONV_FWD, xa, out, pw, k, s, p, bias=b.cuda(), slope=0.0, epilogue=ops.EPI_LEAKY, round_tf32=True)
    torch.cuda.synchronize()
    err = rel(out.to_nchw(), ref)
    assert err < TOL, err


@pytest.mark.parametrize("case", [c for c in DECONV_CASES if c[5] != 8])
def test_deconv_dgrad_wgrad(case):
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p, op = case
    x = _mk(n, cin, h, w, 1).requires_grad_(True)
    wt = (_mk(cin, cout, k, k, 2) * (1.0 / (cout * k * k) ** 0.5)).requires_grad_(True)
    y = F.conv_transpose2d(x, wt, None, stride=s, padding=p, output_padding=op)
    dy = _mk(*y.shape, 7)
    y.backward(dy)
    dya, xa = _act(dy, ops), _act(x.detach(), ops)
    dx = ops.Act.empty(n, h, w, cin)
    pw = ops.pack_weights(wt.detach().cuda(), ops.DECONV_DGRAD, round_tf32=True)
    ops.conv2d(ops.DECONV_DGRAD, dya, dx, pw, k, s, p, round_tf32=True)
    dw = ops.alloc_packed(ops.DECONV_FWD, (cin, cout, k, k))
    ops.wgrad(ops.DECONV_FWD, xa, dya, dw, k, s, p)
    got = torch.zeros((cin, cout, k, k), device="cuda")
    ops.unpack_weights(dw, got, ops.DECONV_FWD)
    torch.cuda.synchronize()
    e1, e2 = rel(dx.to_nchw(), x.grad), rel(got, wt.grad)
    assert e1 < TOL and e2 < TOL, (e1, e2)


def test_epilogue_residual_and_mask():
    """ResidualBlock tail (ModificationLayer.py:300-302): act(conv(h) + x), and the fused activation backward."""
    from tpgan_b200 import ops
    n, c, h, w, k = 2, 80, 64, 64, 5
    hmid, x = _mk(n, c, h, w, 1), _mk(n, c, h, w, 2)
    wt = _mk(c, c, k, k, 3) * (1.0 / (c * k * k) ** 0.5)
    b = _mk(1, c, 1, 1, 4).flatten()
    ref = F.leaky_relu(F.conv2d(hmid, wt, b, padding=2) + x, 0.01)
    out = ops.Act.empty(n, h, w, c)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
    ops.conv2d(ops.CONV_FWD, _act(hmid, ops), out, pw, k, 1, 2, bias=b.cuda(), add1=_act(x, ops), slope=0.01,
               epilogue=ops.EPI_LEAKY, round_tf32=True)
    torch.cuda.synchronize()
    assert rel(out.to_nchw(), ref) < TOL
    # dgrad with two addends and a per-channel mask
    dy, a1, a2, msrc = _mk(n, c, h, w, 5), _mk(n, c, h, w, 6), _mk(n, c, h, w, 7), _mk(n, c, h, w, 8)
    slopes = torch.where(torch.arange(c) % 3 == 0, 0.0, 0.01)
    slopes[5] = 1.0
    g = torch.nn.grad.conv2d_input((n, c, h, w), wt, dy, padding=2) + a1 + a2
    refm = torch.where(msrc > 0, g, g * slopes.view(1, c, 1, 1))
    dx = ops.Act.empty(n, h, w, c)
    pwd = ops.pack_weights(wt.cuda(), ops.CONV_DGRAD, round_tf32=True)
    ops.conv2d(ops.CONV_DGRAD, _act(dy, ops), dx, pwd, k, 1, 2, add1=_act(a1, ops), add2=_act(a2, ops),
               mask=_act(msrc, ops), slopes=slopes.cuda(), epilogue=ops.EPI_MASK, round_tf32=True)
    torch.cuda.synchronize()
    assert rel(dx.to_nchw(), refm) < TOL


def test_concat_slice_io():
    """Reads a channel slice of a wide buffer and writes into a slice of another (torch.cat elimination)."""
    from tpgan_b200 import ops
    n, h, w = 2, 32, 32
    wide_in = ops.Act.empty(n, h, w, 160)
    x = _mk(n, 128, h, w, 1)
    wide_in.slice(32, 128).from_nchw(x.cuda())
    wt = _mk(64, 128, 3, 3, 2) * 0.03
    ref = F.conv2d(x, wt, None, padding=1)
    wide_out = ops.Act.empty(n, h, w, 208)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
    ops.conv2d(ops.CONV_FWD, wide_in.slice(32, 128), wide_out.slice(140, 64), pw, 3, 1, 1, round_tf32=True)
    torch.cuda.synchronize()
    assert rel(wide_out.slice(140, 64).to_nchw(), ref) < TOL
    assert float(wide_out.slice(0, 140).to_nchw().abs().max()) == 0.0
    assert float(wide_out.slice(204, 4).to_nchw().abs().max()) == 0.0


def test_grouped_local_pathway_shapes():
    """Four problems with the local-pathway patch sizes in one launch (D_and_G_model.py:390-393)."""
    from tpgan_b200 import ops
    shapes = [(40, 40), (40, 40), (32, 40), (32, 48)]
    n, cin, cout = 3, 64, 64
    args, refs, outs = [], [], []
    keep = []
    for i, (h, w) in enumerate(shapes):
        x = _mk(n, cin, h, w, 10 + i)
        wt = _mk(cout, cin, 3, 3, 20 + i) * 0.04
        b = _mk(1, cout, 1, 1, 30 + i).flatten()
        refs.append(F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.01))
        out = ops.Act.empty(n, h, w, cout)
        pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
        xa, bc = _act(x, ops), b.cuda()
        keep += [xa, pw, bc]
        args.append(ops.conv_args(ops.CONV_FWD, xa, out, pw, 3, 1, 1, bias=bc, slope=0.01, epilogue=ops.EPI_LEAKY,
                                  round_tf32=True))
        outs.append(out)
    ops.conv2d_grouped(args)
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert rel(o.to_nchw(), r) < TOL


FLAT_CASES = [
    # n, cin, cout, h, w, k  (stride 1, "same" padding): the flat-slab kernel (csrc/flatconv.cu) in forced mode
    (5, 64, 64, 40, 40, 3),      # local conv0.x: T = 2 row units
    (5, 128, 128, 20, 20, 3),    # local conv1.x
    (5, 128, 128, 16, 24, 3),    # mouth patch at half resolution
    (6, 256, 256, 10, 10, 3),    # local conv2.x: one image per unit
    (7, 256, 256, 8, 10, 3),     # nose patch at quarter resolution
    (8, 96, 64, 8, 8, 3),        # whole-image units: four images per three-tile unit
    (7, 96, 64, 8, 8, 3),        # ... with a ragged last unit
    (3, 75, 40, 24, 24, 5),      # 5x5 taps, partial last K chunk, ragged channel count
    (2, 320, 512, 12, 12, 3),    # two N tiles
    (5, 3, 64, 40, 40, 3),       # one partial K chunk
    (2, 64, 3, 33, 17, 3),       # ragged map, 3 output channels
    (2, 64, 64, 64, 64, 5),      # global-pathway conv1.x: four-tile units, 25 taps, two slab slots
    (2, 80, 80, 64, 64, 5),
    (2, 128, 128, 32, 32, 3),    # global-pathway conv2.x
    (32, 64, 64, 40, 40, 3),     # production batch: several units per CTA (ring / accumulator phases wrap)
    (32, 64, 64, 64, 64, 5),
    (32, 80, 80, 64, 64, 5),
    (32, 128, 128, 32, 32, 3),
]


@pytest.mark.parametrize("case", FLAT_CASES)
def test_flatconv_fwd_dgrad(case, monkeypatch):
    from tpgan_b200 import ops, _lib
    monkeypatch.setenv("TPGAN_FLATCONV", "2")
    n, cin, cout, h, w, k = case
    p = (k - 1) // 2
    x = _mk(n, cin, h, w, 1)
    wt = _mk(cout, cin, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)
    b = _mk(1, cout, 1, 1, 3).flatten()
    res = _mk(n, cout, h, w, 4)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=p) + res, 0.01)
    out = ops.Act.empty(n, h, w, cout)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
    ops.conv2d(ops.CONV_FWD, _act(x, ops), out, pw, k, 1, p, bias=b.cuda(), add1=_act(res, ops), slope=0.01,
               epilogue=ops.EPI_LEAKY, round_tf32=True)
    assert _lib.last_conv_kernel() == "flatconv"
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    assert rel(out.to_nchw(), ref) < TOL
    dy, a1, msrc = _mk(n, cout, h, w, 5), _mk(n, cin, h, w, 6), _mk(n, cin, h, w, 7)
    g = torch.nn.grad.conv2d_input((n, cin, h, w), wt, dy, padding=p) + a1
    refm = torch.where(msrc > 0, g, g * 0.01)
    dx = ops.Act.empty(n, h, w, cin)
    pwd = ops.pack_weights(wt.cuda(), ops.CONV_DGRAD, round_tf32=True)
    ops.conv2d(ops.CONV_DGRAD, _act(dy, ops), dx, pwd, k, 1, p, add1=_act(a1, ops), mask=_act(msrc, ops), slope=0.01,
               epilogue=ops.EPI_MASK, round_tf32=True)
    assert _lib.last_conv_kernel() == "flatconv"
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    assert rel(dx.to_nchw(), refm) < TOL


@pytest.mark.parametrize("div,cin,cout", [(1, 64, 64), (2, 128, 128), (4, 256, 256)])
def test_flatconv_grouped_matches_tapgemm(div, cin, cout, monkeypatch):
    """The four local-pathway patches of one layer in one flat-slab launch: same results (to tf32 accumulation order) as the
    multi-tap GEMM kernel and as ATen."""
    from tpgan_b200 import ops, _lib
    shapes = [(40 // div, 40 // div), (40 // div, 40 // div), (32 // div, 40 // div), (32 // div, 48 // div)]
    n = 5
    args, refs, keep = [], [], []
    for i, (h, w) in enumerate(shapes):
        x = _mk(n, cin, h, w, 10 + i)
        wt = _mk(cout, cin, 3, 3, 20 + i) * (1.0 / (cin * 9) ** 0.5)
        b = _mk(1, cout, 1, 1, 30 + i).flatten()
        refs.append(F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.01))
        xa, bc, pw = _act(x, ops), b.cuda(), ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
        keep += [xa, bc, pw]
        args.append((xa, pw, bc, h, w))
    outs = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("TPGAN_FLATCONV", mode)
        o = [ops.Act.empty(n, h, w, cout) for (_, _, _, h, w) in args]
        l0 = _lib.launch_count()
        ops.conv2d_grouped([ops.conv_args(ops.CONV_FWD, xa, oo, pw, 3, 1, 1, bias=bc, slope=0.01, epilogue=ops.EPI_LEAKY,
                                          round_tf32=True) for (xa, pw, bc, _, _), oo in zip(args, o)])
        assert _lib.launch_count() - l0 == 1
        assert _lib.last_conv_kernel() == ("flatconv" if mode == "2" else "tapgemm")
        torch.cuda.synchronize()
        assert _lib.kernel_status() == 0
        outs[mode] = o
    for a, b_, r in zip(outs["0"], outs["2"], refs):
        assert rel(b_.to_nchw(), r) < TOL
        assert rel(b_.to_nchw(), a.to_nchw()) < 2e-4


@pytest.mark.parametrize("case", [(32, 512, 512, 8, 8, 8), (32, 8192, 256, 1, 1, 1), (5, 4096, 48, 1, 1, 1), (3, 64, 1000, 4, 4, 4)])
def test_linear_split_k(case):
    """Linear layers as GEMM-like launches (rows = images): fc1 is an 8x8 'conv' of 64 taps on the 8x8x512 map
    (D_and_G_model.py:289), the identity network's FC a one-tap product over 8192 features.  Few tiles and a long reduction:
    the library splits the taps / the K chunks over the SMs; the CTA arriving last adds the parked partial sums in a fixed
    order and runs the fused epilogue (api.cu, tapgemm.cu)."""
    from tpgan_b200 import ops, _lib
    n, cin, cout, h, w, k = case
    x = _mk(n, cin, h, w, 1)
    wt = _mk(cout, cin, k, k, 2) * (1.0 / (cin * k * k) ** 0.5)
    b = _mk(1, cout, 1, 1, 3).flatten()
    ref = F.leaky_relu(F.conv2d(x, wt, b), 0.01)
    pw = ops.pack_weights(wt.cuda(), ops.CONV_FWD, round_tf32=True)
    xa, bc = _act(x, ops), b.cuda()
    outs = []
    for _ in range(3):          # repeated launches: the arrival counters reset themselves, the result is bit-reproducible
        out = ops.Act.empty(n, 1, 1, cout)
        out.buf.fill_(7.0)
        ops.conv2d(ops.CONV_FWD, xa, out, pw, k, 1, 0, bias=bc, slope=0.01, epilogue=ops.EPI_LEAKY, round_tf32=True)
        outs.append(out.to_nchw())
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    assert rel(outs[0], ref) < TOL
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("case", [(2, 3, 64, 128, 128, 7, 1, 3), (3, 3, 64, 128, 128, 3, 2, 1), (4, 3, 48, 40, 56, 5, 1, 2), (2, 4, 64, 33, 47, 3, 1, 1)])
def test_wgrad_im2col_rgb_slice(case):
    """Weight gradients of the RGB input layers (generator conv0.0: 3->64 k7, critic layer 0: 3->64 k3 stride 2) take the
    im2col-by-TMA path: a zero-padded float4 copy of the input read through tensor maps with overlapping 8-pixel windows.
    Here the input is a channel SLICE of a wider concat buffer (as I128 is in the generator: D_and_G_model.py:312)."""
    from tpgan_b200 import ops
    n, cin, cout, h, w, k, s, p = case
    x = _mk(n, cin, h, w, 1)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dy = _mk(n, cout, ho, wo, 5)
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=p)
    wide = ops.Act.empty(n, h, w, 76)
    wide.buf.uniform_(-1, 1)
    xa = wide.slice(72, cin) if cin <= 4 else None
    xa.from_nchw(x.cuda(), round_tf32=True)
    dw = ops.alloc_packed(ops.CONV_FWD, (cout, cin, k, k))
    for _ in range(2):     # the second launch takes the next slot of the padded-copy ring and accumulates
        ops.wgrad(ops.CONV_FWD, xa, _act(dy, ops), dw, k, s, p)
    got = torch.zeros((cout, cin, k, k), device="cuda")
    ops.unpack_weights(dw, got, ops.CONV_FWD)
    torch.cuda.synchronize()
    assert rel(got, 2 * ref) < TOL


GROUPED_WGRAD_CASES = [
    # (cin, cout, k, stride, pad, kind, patch sizes of the four pathways at this depth) - D_and_G_model.py:390-393 patches
    # 40x40 / 40x40 / 32x40 / 32x48 after 0..3 stride-2 convs.  10x10 and 5x5 boxes do not fill their last 8-pixel K step while
    # the nose / mouth boxes do: the launch mixes box shapes and must re-zero the unwritten K rows every stage.
    (64, 64, 3, 1, 1, "conv", [(40, 40), (40, 40), (32, 40), (32, 48)]),
    (128, 128, 3, 1, 1, "conv", [(20, 20), (20, 20), (16, 20), (16, 24)]),
    (256, 256, 3, 1, 1, "conv", [(10, 10), (10, 10), (8, 10), (8, 12)]),
    (512, 512, 3, 1, 1, "conv", [(5, 5), (5, 5), (4, 5), (4, 6)]),
    (256, 512, 3, 2, 1, "conv", [(10, 10), (10, 10), (8, 10), (8, 12)]),
    (512, 256, 3, 2, 1, "deconv", [(5, 5), (5, 5), (4, 5), (4, 6)]),
]


@pytest.mark.parametrize("case", GROUPED_WGRAD_CASES)
def test_grouped_local_pathway_wgrad(case):
    """The four pathways' weight gradients of one layer in ONE launch (mixed pixel-box shapes included), each equal to ATen's."""
    from tpgan_b200 import ops, _lib
    cin, cout, k, s, p, kind, shapes = case
    n = 5
    args, refs, dws, keep = [], [], [], []
    for i, (h, w) in enumerate(shapes):
        x = _mk(n, cin, h, w, 40 + i)
        if kind == "conv":
            ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
            dy = _mk(n, cout, ho, wo, 50 + i)
            refs.append(torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=p))
            wshape, akind = (cout, cin, k, k), ops.CONV_FWD
        else:
            ho, wo = (h - 1) * s - 2 * p + k + 1, (w - 1) * s - 2 * p + k + 1
            dy = _mk(n, cout, ho, wo, 50 + i)
            wt = torch.zeros(cin, cout, k, k, requires_grad=True)
            F.conv_transpose2d(x, wt, stride=s, padding=p, output_padding=1).backward(dy)
            refs.append(wt.grad)
            wshape, akind = (cin, cout, k, k), ops.DECONV_FWD
        xa, dya = _act(x, ops), _act(dy, ops)
        dw = ops.alloc_packed(akind, wshape)
        keep += [xa, dya]
        dws.append((dw, wshape, akind))
        args.append(ops.wgrad_args(akind, xa, dya, dw, k, s, p, accumulate=False))
    l0 = _lib.launch_count()
    ops.wgrad_grouped(args)
    assert _lib.launch_count() - l0 == 1, "grouped weight gradients must be one launch"
    for (dw, wshape, akind), ref in zip(dws, refs):
        got = torch.zeros(wshape, device="cuda")
        ops.unpack_weights(dw, got, akind)
        torch.cuda.synchronize()
        err = rel(got, ref)
        assert err < TOL, err
    assert _lib.kernel_status() == 0


# ---- full-size (BASELINE config 1, batch 32) size-independent property: the three kernels of a layer are adjoint.
#   <conv(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(x, dy)>
# With tf32-representable x, w, dy every product is exact in the tensor cores, so the three inner products differ only
# by fp32 accumulation order and by the tf32 rounding of the stored conv / dgrad outputs (relative 2^-11 per element,
# averaging out over the sum): 1e-3 relative to the norm bound ||y||*||dy|| is generous and still catches a single
# mis-routed tap, tile, channel chunk or K block.  Shapes: SURVEY.md T1 (rowconv / rowstack / tapgemm / wgrad plain,
# slab and swapped modes at their production sizes).
ADJOINT_CASES = [
    # cin, cout, k, stride, pad, H=W
    (206, 206, 5, 1, 2, 128),
    (75, 75, 7, 1, 3, 128),
    (64, 64, 7, 1, 3, 128),
    (206, 64, 5, 1, 2, 128),
    (3, 64, 7, 1, 3, 128),
    (64, 3, 3, 1, 1, 128),
    (64, 64, 5, 2, 2, 128),
    (208, 208, 3, 1, 1, 64),
    (416, 416, 3, 1, 1, 32),
    (768, 768, 3, 1, 1, 16),
    (512, 512, 3, 1, 1, 8),
]


@pytest.mark.parametrize("case", ADJOINT_CASES)
def test_full_size_adjoint_identities(case):
    from oracle.model_port import tf32_rna
    from tpgan_b200 import ops
    cin, cout, k, s, p, H = case
    B = 32
    Ho = (H + 2 * p - k) // s + 1
    g = torch.Generator(device="cuda").manual_seed(123)

    def rnd(*shape, scale=1.0):
        return tf32_rna((torch.rand(*shape, device="cuda", generator=g) * 2 - 1) * scale)
    x = ops.Act.empty(B, H, H, cin)
    x.buf.copy_(rnd(*x.buf.shape))
    dy = ops.Act.empty(B, Ho, Ho, cout)
    dy.buf.copy_(rnd(*dy.buf.shape))
    if cin % 4:
        x.buf[..., cin:] = 0
    if cout % 4:
        dy.buf[..., cout:] = 0
    w = rnd(cout, cin, k, k, scale=0.05)
    y, dx = ops.Act.empty(B, Ho, Ho, cout), ops.Act.empty(B, H, H, cin)
    wf, wd = ops.pack_weights(w, ops.CONV_FWD), ops.pack_weights(w, ops.CONV_DGRAD)
    dw = ops.alloc_packed(ops.CONV_FWD, tuple(w.shape))
    ops.conv2d(ops.CONV_FWD, x, y, wf, k, s, p, round_tf32=False)
    ops.conv2d(ops.CONV_DGRAD, dy, dx, wd, k, s, p, round_tf32=False)
    ops.wgrad(ops.CONV_FWD, x, dy, dw, k, s, p, accumulate=False)
    dwr = torch.zeros_like(w)
    ops.unpack_weights(dw, dwr, ops.CONV_FWD)
    torch.cuda.synchronize()
    dot = lambda a, b: float((a.double() * b.double()).sum())
    a1 = dot(y.buf[..., :cout], dy.buf[..., :cout])
    a2 = dot(x.buf[..., :cin], dx.buf[..., :cin])
    a3 = dot(w, dwr)
    bound = float(y.buf[..., :cout].double().norm() * dy.buf[..., :cout].double().norm())
    assert abs(a1 - a2) <= 1e-3 * bound and abs(a1 - a3) <= 1e-3 * bound, (a1, a2, a3, bound)
    assert abs(a1) > 0 and bound > 0
