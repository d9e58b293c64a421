"""Thin Python operators over the C ABI (include/tpgan_b200.h).

PyTorch is used only for device memory and streams: every function here takes torch CUDA tensors (or `Act` channel
views of NHWC buffers), turns them into raw pointers and calls libtpgan_b200.so.  Nothing computes in ATen.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import (CONV_DGRAD, CONV_FWD, DECONV_DGRAD, DECONV_FWD, DTYPE_BF16, DTYPE_TF32, EPI_LEAKY, EPI_LINEAR, EPI_MASK,
                   NULL_VIEW, ConvArgs, View, WgradArgs)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.dtype in (torch.float32, torch.int32, torch.uint8, torch.bfloat16), (t.device, t.dtype)
    return t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------------------------- zeroed memory
class Arena:
    """Bump allocator over a few large zero-initialised chunks.  A trainer allocates ~2000 buffers (activations, gradient
    buffers, packed weights); as individual torch.zeros calls that is ~2000 fill launches before the first kernel of this
    library runs.  Chunks are ordinary tensors and every allocation is a view of one, so lifetime is plain reference
    counting: a chunk is freed when the last buffer carved out of it dies."""

    def __init__(self, device, chunk_bytes: int = 256 << 20, shadow: bool = False):
        """shadow=True (bf16 operand mode): every chunk has a bf16 twin of half its size; the twin of an fp32 buffer sits at
        the SAME ELEMENT OFFSET in it, so any view of the buffer (channel slice, batch slice, reshape) finds its bf16 operand
        copy by pointer arithmetic with unchanged element strides (Act.view16)."""
        self.device, self.chunk_bytes, self.shadow = torch.device(device), chunk_bytes, shadow
        self.chunk: Optional[torch.Tensor] = None
        self.used = 0
        self.total = 0
        self.twins: list = []     # (fp32 chunk base pointer, bytes, bf16 twin, chunk) of every shadowed chunk
        if shadow:
            import weakref
            _SHADOWS.append(weakref.ref(self))

    def alloc(self, shape, dtype=torch.float32) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = round_up(max(n, 1) * torch.empty((), dtype=dtype).element_size(), 256)   # keeps TMA / vector alignment
        if self.chunk is None or self.used + nbytes > self.chunk.numel():
            self.chunk = torch.zeros(max(self.chunk_bytes, nbytes), dtype=torch.uint8, device=self.device)
            self.used = 0
            if self.shadow:
                twin = torch.zeros(self.chunk.numel() // 2, dtype=torch.uint8, device=self.device)
                self.twins.append((self.chunk.data_ptr(), self.chunk.numel(), twin, self.chunk))
        o = self.used
        self.used += nbytes
        self.total += nbytes
        return self.chunk[o:o + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)


_ARENA: List[Arena] = []
_SHADOWS: list = []     # weak references to the shadowed arenas alive (a bf16 trainer owns one)


def _find_twin(ptr: int):
    dead = False
    for ref in _SHADOWS:
        a = ref()
        if a is None:
            dead = True
            continue
        for base, nbytes, twin, _ in a.twins:
            if base <= ptr < base + nbytes:
                return base, twin
    if dead:
        _SHADOWS[:] = [r for r in _SHADOWS if r() is not None]
    raise RuntimeError("tpgan_b200: buffer has no bf16 twin (allocate it inside use_arena(Arena(shadow=True)))")


def shadow_ptr(ptr: int) -> int:
    """Address of the bf16 twin of the fp32 element at device address `ptr` (must lie in a shadowed arena chunk)."""
    base, twin = _find_twin(ptr)
    return twin.data_ptr() + (ptr - base) // 2


class use_arena:
    """with use_arena(a): every ops.zeros() / Act.empty() on a's device inside is carved out of `a`."""

    def __init__(self, arena: Arena):
        self.arena = arena

    def __enter__(self):
        _ARENA.append(self.arena)
        return self.arena

    def __exit__(self, *exc):
        _ARENA.pop()
        return False


def _same_device(a: torch.device, b: torch.device) -> bool:
    """torch.device("cuda") and torch.device("cuda", i) name the same device when i is the current one."""
    if a.type != b.type:
        return False
    if a.type != "cuda" or a.index == b.index:
        return True
    cur = torch.cuda.current_device()
    return (cur if a.index is None else a.index) == (cur if b.index is None else b.index)


def zeros(shape, dtype=torch.float32, device="cuda") -> torch.Tensor:
    """Zero-initialised device tensor; inside a use_arena() block a view of the arena (no fill launch)."""
    if isinstance(shape, int):
        shape = (shape,)
    if _ARENA and _same_device(_ARENA[-1].device, torch.device(device)):
        return _ARENA[-1].alloc(tuple(shape), dtype)
    return torch.zeros(tuple(shape), dtype=dtype, device=device)


class Act:
    """Channel slice [c0, c0+c) of an NHWC fp32 buffer `buf` of shape (N, H, W, Cs)."""

    __slots__ = ("buf", "c0", "c")

    def __init__(self, buf: torch.Tensor, c0: int = 0, c: Optional[int] = None):
        assert buf.dim() == 4 and buf.dtype == torch.float32 and buf.is_contiguous(), (buf.shape, buf.dtype)
        self.buf = buf
        self.c0 = c0
        self.c = buf.shape[3] - c0 if c is None else c
        assert 0 <= c0 and c0 + self.c <= buf.shape[3]

    @staticmethod
    def empty(n: int, h: int, w: int, c: int, device="cuda", zero: bool = True) -> "Act":
        # channel stride: 16-byte pixels for fp32 (TMA rows); in a shadowed (bf16) arena 8 channels, so that the bf16 twin's
        # pixel stride is a multiple of 16 bytes as well
        cs = round_up(c, 8 if (_ARENA and _ARENA[-1].shadow) else 4)
        buf = zeros((n, h, w, cs), torch.float32, device) if (zero or _ARENA) else \
            torch.empty((n, h, w, cs), dtype=torch.float32, device=device)
        return Act(buf, 0, c)

    @property
    def n(self):
        return self.buf.shape[0]

    @property
    def h(self):
        return self.buf.shape[1]

    @property
    def w(self):
        return self.buf.shape[2]

    def slice(self, c0: int, c: int) -> "Act":
        return Act(self.buf, self.c0 + c0, c)

    def view(self) -> View:
        n, h, w, cs = self.buf.shape
        return View(self.buf.data_ptr() + 4 * self.c0, h * w * cs, w * cs, cs, n, h, w, self.c)

    def view16(self) -> View:
        """The bf16 twin of this view (same element strides; ptr addresses 2-byte elements)."""
        n, h, w, cs = self.buf.shape
        return View(shadow_ptr(self.buf.data_ptr() + 4 * self.c0), h * w * cs, w * cs, cs, n, h, w, self.c)

    def twin(self) -> torch.Tensor:
        """The bf16 twin as a torch tensor of this view's logical shape (N, H, W, c) - tests / debugging."""
        n, h, w, cs = self.buf.shape
        p = self.buf.data_ptr() + 4 * self.c0
        base, tw = _find_twin(p)
        return torch.as_strided(tw.view(torch.bfloat16), (n, h, w, self.c), (h * w * cs, w * cs, cs, 1), (p - base) // 4)

    def like(self, zero: bool = True) -> "Act":
        return Act.empty(self.n, self.h, self.w, self.c, self.buf.device, zero)

    def to_nchw(self) -> torch.Tensor:
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.buf.device)
        _lib.check(_lib.load().tpgan_nhwc_to_nchw(self.view(), out.data_ptr(), _stream()), "nhwc_to_nchw")
        return out

    def from_nchw(self, t: torch.Tensor, round_tf32: bool = False) -> "Act":
        t = t.contiguous()
        assert tuple(t.shape) == (self.n, self.c, self.h, self.w), (t.shape, (self.n, self.c, self.h, self.w))
        _lib.check(_lib.load().tpgan_nchw_to_nhwc(t.data_ptr(), self.view(), int(round_tf32), _stream()), "nchw_to_nhwc")
        return self


def _v(a: Optional[Act]) -> View:
    return NULL_VIEW if a is None else a.view()


@dataclass
class Packed:
    """Tensor-core weight layout [taps+1][rows_pad][k_pad] (see tpgan_pack_weights)."""
    data: torch.Tensor
    taps: int
    rows: int
    k: int
    rows_pad: int
    k_pad: int


def pack_geometry(kind: int, w_shape: Sequence[int]):
    """(rows, k, taps, row_stride, k_stride) of the reference weight for the GEMM of `kind`."""
    d0, d1, kh, kw = w_shape
    taps = kh * kw
    if kind == CONV_FWD:      # weight (Cout, Cin, kh, kw): rows = Cout, k = Cin
        return d0, d1, taps, d1 * taps, taps
    if kind == CONV_DGRAD:    # rows = Cin, k = Cout
        return d1, d0, taps, taps, d1 * taps
    if kind == DECONV_FWD:    # weight (Cin, Cout, kh, kw): rows = Cout, k = Cin
        return d1, d0, taps, taps, d1 * taps
    if kind == DECONV_DGRAD:  # rows = Cin, k = Cout
        return d0, d1, taps, d1 * taps, taps
    raise ValueError(kind)


def alloc_packed(kind: int, w_shape: Sequence[int], rows_int: Optional[int] = None, k_int: Optional[int] = None,
                 device="cuda") -> Packed:
    rows, k, taps, _, _ = pack_geometry(kind, w_shape)
    rows = rows if rows_int is None else rows_int
    k = k if k_int is None else k_int
    rows_pad, k_pad = round_up(rows, 16), round_up(k, 32)
    data = zeros((taps + 1, rows_pad, k_pad), torch.float32, device)
    return Packed(data, taps, rows, k, rows_pad, k_pad)


def pack_weights(w: torch.Tensor, kind: int, out: Optional[Packed] = None, row_map: Optional[torch.Tensor] = None,
                 k_map: Optional[torch.Tensor] = None, round_tf32=True) -> Packed:
    """Reference-layout weight -> tensor-core layout for the GEMM of `kind`.  row_map/k_map (int32, device) give for each
    internal row / k index the reference channel (-1 = padding)."""
    w = w.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    rows, k, taps, rs, ks = pack_geometry(kind, w.shape)
    rows_int = rows if row_map is None else row_map.numel()
    k_int = k if k_map is None else k_map.numel()
    if out is None:
        out = alloc_packed(kind, w.shape, rows_int, k_int, w.device)
    _lib.check(_lib.load().tpgan_pack_weights(w.data_ptr(), out.data.data_ptr(), taps, rows_int, k_int, out.rows_pad,
                                              out.k_pad, rs, ks, _ptr(row_map), _ptr(k_map), int(round_tf32), _stream()),
               "pack_weights")
    return out


def unpack_weights(packed: Packed, w_grad: torch.Tensor, kind: int = CONV_FWD, row_map=None, k_map=None,
                   accumulate: bool = False) -> None:
    """Packed (forward-layout) weight gradient -> reference layout gradient tensor."""
    assert w_grad.is_contiguous()
    rows, k, taps, rs, ks = pack_geometry(kind, w_grad.shape)
    rows_int = rows if row_map is None else row_map.numel()
    k_int = k if k_map is None else k_map.numel()
    _lib.check(_lib.load().tpgan_unpack_weights(packed.data.data_ptr(), w_grad.data_ptr(), taps, rows_int, k_int,
                                                packed.rows_pad, packed.k_pad, rs, ks, _ptr(row_map), _ptr(k_map),
                                                int(accumulate), _stream()), "unpack_weights")


def transpose_packed(src: Packed, dst: Packed) -> None:
    """dst[t][kk][r] = src[t][r][kk] over the valid region (forward packing <-> input-gradient packing)."""
    assert src.taps == dst.taps and src.rows == dst.k and src.k == dst.rows
    _lib.check(_lib.load().tpgan_transpose_packed(src.data.data_ptr(), dst.data.data_ptr(), src.taps, src.rows, src.k,
                                                  src.rows_pad, src.k_pad, dst.rows_pad, dst.k_pad, _stream()),
               "transpose_packed")


def conv_args(kind: int, x: Act, out: Act, w: Packed, k: int, stride: int, pad: int, bias: Optional[torch.Tensor] = None,
              add1: Optional[Act] = None, add2: Optional[Act] = None, mask: Optional[Act] = None,
              slopes: Optional[torch.Tensor] = None, slope: float = 0.0, epilogue: int = EPI_LINEAR,
              round_tf32: bool = True, bf16: bool = False, out16: bool = True, out32: bool = True,
              x_lo: Optional[Act] = None, w_lo: Optional[Packed] = None) -> ConvArgs:
    """bf16=True: x is read through its bf16 twin, `w` is a bf16 packing (Packed.data of dtype bfloat16), the result is
    written to out (fp32, unless out32=False) and to out's bf16 twin (unless out16=False).
    x_lo / w_lo (tf32 only): residual parts of the operands - the launch computes the fp32-accurate three-term product."""
    if not bf16:
        assert (x_lo is None) == (w_lo is None)
        return ConvArgs(kind, k, k, stride, pad, x.view(), out.view(), w.data.data_ptr(), w.rows_pad, w.k_pad, _ptr(bias),
                        _v(add1), _v(add2), _v(mask), _ptr(slopes), float(slope), epilogue, int(round_tf32), DTYPE_TF32,
                        NULL_VIEW, _v(x_lo), None if w_lo is None else w_lo.data.data_ptr())
    assert w.data.dtype == torch.bfloat16 and (out16 or out32)
    ov = out.view()
    if not out32:
        ov.ptr = None
    return ConvArgs(kind, k, k, stride, pad, x.view16(), ov, w.data.data_ptr(), w.rows_pad, w.k_pad, _ptr(bias),
                    _v(add1), _v(add2), _v(mask), _ptr(slopes), float(slope), epilogue, 0, DTYPE_BF16,
                    out.view16() if out16 else NULL_VIEW, NULL_VIEW, None)


def conv2d_grouped(args: List[ConvArgs]) -> None:
    arr = (ConvArgs * len(args))(*args)
    _lib.check(_lib.load().tpgan_conv2d(arr, len(args), _stream()), "conv2d")


def conv2d(*a, **kw) -> None:
    conv2d_grouped([conv_args(*a, **kw)])


def wgrad_args(kind: int, x: Act, dy: Act, dw: Packed, k: int, stride: int, pad: int, accumulate: bool = True,
               bf16: bool = False) -> WgradArgs:
    if bf16:    # operands = the bf16 twins of x and dy; dw stays the fp32 forward-packed accumulator
        return WgradArgs(kind, k, k, stride, pad, x.view16(), dy.view16(), dw.data.data_ptr(), dw.rows_pad, dw.k_pad,
                         int(accumulate), DTYPE_BF16)
    return WgradArgs(kind, k, k, stride, pad, x.view(), dy.view(), dw.data.data_ptr(), dw.rows_pad, dw.k_pad,
                     int(accumulate), DTYPE_TF32)


def cast_bf16(a: Act) -> None:
    """bf16 twin of `a` <- rne(a): operand copy of a tensor that no tensor-core epilogue wrote."""
    _lib.check(_lib.load().tpgan_cast_bf16(a.view(), a.view16(), _stream()), "cast_bf16")


def alloc_packed16(pk: Packed) -> Packed:
    """bf16 operand copy of a packing: [taps+1][rows_pad][k_pad16], k_pad16 = k_pad rounded up to 64."""
    k16 = round_up(pk.k_pad, 64)
    data = zeros((pk.data.shape[0], pk.rows_pad, k16), torch.bfloat16, pk.data.device)
    return Packed(data, pk.taps, pk.rows, pk.k, pk.rows_pad, k16)


def cast_job(src: Packed, dst: Packed) -> "_lib.CastJob":
    assert dst.data.dtype == torch.bfloat16 and src.rows_pad == dst.rows_pad and src.data.shape[0] == dst.data.shape[0]
    return _lib.CastJob(src.data.data_ptr(), dst.data.data_ptr(), src.data.shape[0] * src.rows_pad, src.k_pad, dst.k_pad, 0, 0)


_CAST_TABLES: dict = {}


def cast_packed(src: Packed, dst: Packed) -> None:
    """dst (bf16 packing) <- src (fp32 packing).  The one-job table is built once per (src, dst) pair: building it copies
    host memory to the device, which must not happen while a CUDA graph is being captured."""
    key = (src.data.data_ptr(), dst.data.data_ptr())
    tab = _CAST_TABLES.get(key)
    if tab is None:
        tab = _CAST_TABLES[key] = (JobTable("cast", [cast_job(src, dst)], src.data.device), src, dst)
    tab[0].run()


def wgrad_grouped(args: List[WgradArgs]) -> None:
    arr = (WgradArgs * len(args))(*args)
    _lib.check(_lib.load().tpgan_conv2d_wgrad(arr, len(args), _stream()), "conv2d_wgrad")


def wgrad(*a, **kw) -> None:
    wgrad_grouped([wgrad_args(*a, **kw)])


def act_backward(src: Act, mask: Act, dst: Act, slope: float = 0.0, slopes: Optional[torch.Tensor] = None) -> None:
    _lib.check(_lib.load().tpgan_act_backward(src.view(), mask.view(), dst.view(), _ptr(slopes), float(slope), _stream()),
               "act_backward")


def view_copy(src: Act, dst: Act, accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_view_copy(src.view(), dst.view(), int(accumulate), _stream()), "view_copy")


def bias_grad(dy: Act, db: torch.Tensor, accumulate: bool = True) -> None:
    _lib.check(_lib.load().tpgan_bias_grad(dy.view(), db.data_ptr(), int(accumulate), _stream()), "bias_grad")


def reflect_pad(src: Act, dst: Act, left: int, top: int) -> None:
    _lib.check(_lib.load().tpgan_reflect_pad(src.view(), dst.view(), left, top, _stream()), "reflect_pad")


def reflect_pad_backward(dpad: Act, dsrc: Act, left: int, top: int, accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_reflect_pad_backward(dpad.view(), dsrc.view(), left, top, int(accumulate), _stream()),
               "reflect_pad_backward")


def patch_crop(img: Act, landmarks: torch.Tensor, patches: Sequence[Act], boxes: Optional[torch.Tensor] = None,
               fill: float = -1.0) -> None:
    assert landmarks.dtype == torch.float32 and landmarks.is_contiguous() and tuple(landmarks.shape) == (img.n, 5, 2)
    _lib.check(_lib.load().tpgan_patch_crop(img.view(), landmarks.data_ptr(), *[p.view() for p in patches], _ptr(boxes),
                                            float(fill), _stream()), "patch_crop")


def local_fuse(patches: Sequence[Act], out: Act, argmax: Optional[torch.Tensor] = None) -> None:
    _lib.check(_lib.load().tpgan_local_fuse(*[p.view() for p in patches], out.view(), _ptr(argmax), _stream()),
               "local_fuse")


def local_fuse_backward(dout: Act, argmax: torch.Tensor, dpatches: Sequence[Act], accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_local_fuse_backward(dout.view(), argmax.data_ptr(), *[p.view() for p in dpatches],
                                                     int(accumulate), _stream()), "local_fuse_backward")


def maxpool3s2(x: Act, y: Act, argmax: Optional[torch.Tensor]) -> None:
    _lib.check(_lib.load().tpgan_maxpool3s2(x.view(), y.view(), _ptr(argmax), _stream()), "maxpool3s2")


def maxpool3s2_backward(dy: Act, argmax: torch.Tensor, dx: Act, accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_maxpool3s2_backward(dy.view(), argmax.data_ptr(), dx.view(), int(accumulate), _stream()),
               "maxpool3s2_backward")


def avgpool(x: Act, y: Act) -> None:
    _lib.check(_lib.load().tpgan_avgpool(x.view(), y.view(), _stream()), "avgpool")


def avgpool_backward(dy: Act, dx: Act, accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_avgpool_backward(dy.view(), dx.view(), int(accumulate), _stream()), "avgpool_backward")


def image_losses(fake: Act, t128: Act, t64: Act, t32: Act, dfake: Act, coeffs: Sequence[float],
                 sums: torch.Tensor) -> None:
    w = (C.c_float * 8)(*coeffs)
    _lib.check(_lib.load().tpgan_image_losses(fake.view(), t128.view(), t64.view(), t32.view(), dfake.view(), w,
                                              sums.data_ptr(), _stream()), "image_losses")


def l1_loss(a: Act, b: Act, da: Optional[Act], coeff: float, total: torch.Tensor) -> None:
    _lib.check(_lib.load().tpgan_l1_loss(a.view(), b.view(), _v(da), float(coeff), total.data_ptr(), _stream()), "l1_loss")


def maxout2(x: torch.Tensor, y: torch.Tensor) -> None:
    rows, cols = y.shape
    _lib.check(_lib.load().tpgan_maxout2(x.data_ptr(), y.data_ptr(), rows, cols, _stream()), "maxout2")


def maxout2_backward(x: torch.Tensor, dy: torch.Tensor, dx: torch.Tensor) -> None:
    rows, cols = dy.shape
    _lib.check(_lib.load().tpgan_maxout2_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), rows, cols, _stream()),
               "maxout2_backward")


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0) -> None:
    _lib.check(_lib.load().tpgan_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                           beta2, eps, weight_decay, step, grad_scale, _stream()), "adam_step")


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_dev: torch.Tensor, grad_scale=1.0) -> None:
    assert step_dev.dtype == torch.int32 and step_dev.is_cuda
    _lib.check(_lib.load().tpgan_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                               beta2, eps, weight_decay, step_dev.data_ptr(), grad_scale, _stream()),
               "adam_step_dev")


def adam_slice_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_dev: torch.Tensor, grad_scale=1.0,
                   increment: bool = False) -> None:
    """Adam over a slice of the flat buffers (see tpgan_adam_slice_dev): the step count is advanced by the first slice only."""
    assert step_dev.dtype == torch.int32 and step_dev.is_cuda and p.numel() == g.numel()
    _lib.check(_lib.load().tpgan_adam_slice_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                                beta2, eps, weight_decay, step_dev.data_ptr(), grad_scale, int(increment),
                                                _stream()), "adam_slice_dev")


def sample_sqnorm(g: Act, out: torch.Tensor) -> None:
    _lib.check(_lib.load().tpgan_sample_sqnorm(g.view(), out.data_ptr(), _stream()), "sample_sqnorm")


def sample_scale(g: Act, coeff: torch.Tensor, u: Act) -> None:
    _lib.check(_lib.load().tpgan_sample_scale(g.view(), coeff.data_ptr(), u.view(), _stream()), "sample_scale")


def gp_coeff(sqnorm: torch.Tensor, coeff: torch.Tensor, scale: float, gp_sum: Optional[torch.Tensor]) -> None:
    _lib.check(_lib.load().tpgan_gp_coeff(sqnorm.data_ptr(), coeff.data_ptr(), sqnorm.numel(), float(scale), _ptr(gp_sum),
                                          _stream()), "gp_coeff")


def lerp(a: Act, b: Act, alpha: torch.Tensor, out: Act) -> None:
    assert alpha.dtype == torch.float32 and alpha.numel() == a.n
    _lib.check(_lib.load().tpgan_lerp(a.view(), b.view(), alpha.data_ptr(), out.view(), _stream()), "lerp")


def mul(a: Act, b: Act, out: Act) -> None:
    _lib.check(_lib.load().tpgan_mul(a.view(), b.view(), out.view(), _stream()), "mul")


def split_tf32(src: Act, hi: Act, lo: Act) -> None:
    _lib.check(_lib.load().tpgan_split_tf32(src.view(), hi.view(), lo.view(), _stream()), "split_tf32")


def fill(out: Act, value: float) -> None:
    _lib.check(_lib.load().tpgan_fill(out.view(), float(value), _stream()), "fill")


def softmax_ce(logits: Act, labels: torch.Tensor, dlogits: Optional[Act], coeff: float, loss_sum: torch.Tensor) -> None:
    """logits: (B,1,1,C) Act; labels int64 (B,)."""
    assert labels.dtype == torch.int64 and labels.is_cuda and logits.h == 1 and logits.w == 1
    lv = logits.view()
    dptr, dstride = (None, 0) if dlogits is None else (dlogits.view().ptr, dlogits.view().sn)
    _lib.check(_lib.load().tpgan_softmax_ce(lv.ptr, lv.sn, labels.data_ptr(), dptr, dstride, logits.n, logits.c,
                                            float(coeff), loss_sum.data_ptr(), _stream()), "softmax_ce")


# ---------------------------------------------------------------------------------------------------- multi-tensor launches
class JobTable:
    """A device-resident array of job structs + the launch that consumes it (one launch for all layers)."""

    def __init__(self, kind: str, jobs: list, device, unpack: bool = False):
        self.kind, self.n, self.unpack = kind, len(jobs), unpack
        self.blocks, self.max_row = 0, 1
        if not jobs:
            return
        for j in jobs:
            j.block_begin = self.blocks
            if kind == "bias":
                self.blocks += j.pix_blocks * j.cgroups
            elif kind == "cast":
                self.blocks += (j.rows * j.k_pad16 // 4 + 2047) // 2048
            elif kind == "pack":
                self.blocks += j.rows if unpack else j.rows_pad
                # shared-memory floats of one staged row: [k][tap] with an odd tap stride (pack_multi_kernel)
                self.max_row = max(self.max_row, j.row_len // j.taps * (j.taps | 1) + j.row_len % j.taps)
            else:
                self.blocks += j.taps * j.tiles_r * j.tiles_k
        arr = (type(jobs[0]) * len(jobs))(*jobs)
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)

    def run(self):
        if self.n == 0:
            return
        lib = _lib.load()
        if self.kind == "bias":
            _lib.check(lib.tpgan_bias_grad_multi(self.table.data_ptr(), self.n, self.blocks, _stream()), "bias_grad_multi")
        elif self.kind == "cast":
            _lib.check(lib.tpgan_cast_packed_multi(self.table.data_ptr(), self.n, self.blocks, _stream()), "cast_packed_multi")
        elif self.kind == "pack":
            _lib.check(lib.tpgan_pack_multi(self.table.data_ptr(), self.n, self.blocks, self.max_row, int(self.unpack),
                                            _stream()), "pack_multi")
        else:
            _lib.check(lib.tpgan_transpose_multi(self.table.data_ptr(), self.n, self.blocks, _stream()), "transpose_multi")


def bias_job(g: Act, db: torch.Tensor) -> "_lib.BiasJob":
    v = g.view()
    assert v.sh == v.w * v.sw and v.sn == v.h * v.sh and v.sw % 4 == 0 and v.ptr % 16 == 0, "bias job needs a dense view"
    npix = v.n * v.h * v.w
    lanes = 1
    while lanes < 32 and lanes * 4 < v.c:
        lanes *= 2
    ppb = 8 * (32 // lanes)                      # pixels one 8-warp block takes per sweep
    # ~64 sweeps per block (8 loads of 16 B in flight per lane), at most two waves of blocks per tensor
    pix_blocks = max(1, min((npix + 64 * ppb - 1) // (64 * ppb), 296))
    if _lib.deterministic():
        pix_blocks = 1      # one block per channel group: a single atomic add per channel onto the cleared accumulator
    return _lib.BiasJob(v.ptr, db.data_ptr(), npix, v.sw, v.c, 0, pix_blocks, (v.c + 4 * lanes - 1) // (4 * lanes), lanes, 0)


def pack_job(ref: torch.Tensor, packed: Packed, taps: int, row_len: int, row_map, k_map, flag: int) -> "_lib.PackJob":
    p = ref.data_ptr()
    return _lib.PackJob(p, p, packed.data.data_ptr(), _ptr(row_map), _ptr(k_map), row_len, taps, packed.rows, packed.k,
                        packed.rows_pad, packed.k_pad, row_len, flag, 0)


def transpose_job(src: Packed, dst: Packed) -> "_lib.TransposeJob":
    assert src.taps == dst.taps and src.rows == dst.k and src.k == dst.rows
    return _lib.TransposeJob(src.data.data_ptr(), dst.data.data_ptr(), src.taps, src.rows, src.k, src.rows_pad, src.k_pad,
                             dst.rows_pad, dst.k_pad, 0, (src.k + 31) // 32, (src.rows + 31) // 32)


# ---------------------------------------------------------------------------------------------------------- Pretrain path
def dwconv3x3(x: Act, y: Act, w: torch.Tensor, stride: int) -> None:
    """Depthwise 3x3 conv, pad 1 (MobileNetV2.py:110); w = the reference (C,1,3,3) tensor."""
    _lib.check(_lib.load().tpgan_dwconv3x3(x.view(), y.view(), _ptr(w), stride, _stream()), "dwconv3x3")


def dwconv3x3_dgrad(dy: Act, dx: Act, w: torch.Tensor, stride: int, accumulate: bool = False) -> None:
    _lib.check(_lib.load().tpgan_dwconv3x3_dgrad(dy.view(), dx.view(), _ptr(w), stride, int(accumulate), _stream()),
               "dwconv3x3_dgrad")


def dwconv3x3_wgrad(x: Act, dy: Act, dw: torch.Tensor, stride: int) -> None:
    """dw (C,1,3,3) += ... (atomics)."""
    _lib.check(_lib.load().tpgan_dwconv3x3_wgrad(x.view(), dy.view(), _ptr(dw), stride, _stream()), "dwconv3x3_wgrad")


def bn_forward(x: Act, res: Optional[Act], y: Act, gamma, beta, running_mean, running_var, momentum: float, eps: float,
               training: bool, relu6: bool, round_tf32: bool, sums: torch.Tensor, coef: torch.Tensor,
               slope: Optional[float] = None) -> None:
    """nn.BatchNorm2d (+ReLU6 | + (Leaky)ReLU(slope), + residual) forward; sums = float64[2C+1] scratch, coef =
    float32[4C] (kept for backward)."""
    assert sums.dtype == torch.float64 and coef.dtype == torch.float32 and not (relu6 and slope is not None)
    act = 1 if relu6 else (2 if slope is not None else 0)
    _lib.check(_lib.load().tpgan_bn_forward(x.view(), _v(res), y.view(), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                            _ptr(running_var), momentum, eps, int(training), act,
                                            0.0 if slope is None else float(slope), int(round_tf32), sums.data_ptr(),
                                            _ptr(coef), _stream()), "bn_forward")


def bn_backward(dy: Act, x: Act, dx: Act, coef: torch.Tensor, training: bool, relu6: bool, accumulate: bool,
                round_tf32: bool, dsums: torch.Tensor, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor]) -> None:
    assert dsums.dtype == torch.float64
    _lib.check(_lib.load().tpgan_bn_backward(dy.view(), x.view(), dx.view(), _ptr(coef), int(training), int(relu6),
                                             int(accumulate), int(round_tf32), dsums.data_ptr(), _ptr(dgamma), _ptr(dbeta),
                                             _stream()), "bn_backward")


def rows_gather(v: Act, flat: torch.Tensor, row_stride: int, offset: int, reverse: bool = False) -> None:
    """SSDHead view(N,-1,K) + cat(dim=1) (MobileNetV2.py:62-76): v's per-image floats <-> flat[n, offset:offset+h*w*c]."""
    _lib.check(_lib.load().tpgan_rows_gather(v.view(), _ptr(flat), row_stride, offset, int(reverse), _stream()),
               "rows_gather")


def multitask_loss(loc: torch.Tensor, cls: torch.Tensor, truth: torch.Tensor, u: Optional[torch.Tensor], n: int,
                   loc_stride: int, cls_stride: int, k_near: int, img_w: float, img_h: float, alpha: float, beta: float,
                   ratio_nb: float, coeff: float, dloc: Optional[torch.Tensor], dcls: Optional[torch.Tensor],
                   labels: Optional[torch.Tensor], sums: torch.Tensor) -> None:
    """Batched MultiTaskLoss (MobileNetV2.py:342-534); see include/tpgan_b200.h."""
    B = truth.shape[0]
    assert truth.is_contiguous() and truth.numel() == B * 8 and (u is None or (u.is_contiguous() and u.numel() == B * n))
    assert labels is None or (labels.dtype == torch.int32 and labels.numel() == B * n)
    _lib.check(_lib.load().tpgan_multitask_loss(_ptr(loc), _ptr(cls), _ptr(truth), _ptr(u), B, n, loc_stride, cls_stride, 5,
                                                k_near, img_w, img_h, alpha, beta, ratio_nb, coeff, _ptr(dloc), _ptr(dcls),
                                                _ptr(labels), _ptr(sums), _stream()), "multitask_loss")


def sgd_step(p: torch.Tensor, g: torch.Tensor, buf: torch.Tensor, lr_dev: torch.Tensor, momentum: float,
             weight_decay: float, nesterov: bool, grad_scale: float = 1.0) -> None:
    """torch.optim.SGD (UtilityMethods.py:30) over flat buffers; lr read from device memory."""
    _lib.check(_lib.load().tpgan_sgd_step(_ptr(p), _ptr(g), _ptr(buf), p.numel(), _ptr(lr_dev), momentum, weight_decay,
                                          int(nesterov), grad_scale, _stream()), "sgd_step")


def ssd_decode(loc: torch.Tensor, cls: torch.Tensor, n: int, loc_stride: int, cls_stride: int, num_classes: int, top_k: int,
               confidence_threshold: float, nms_distance: float, count: torch.Tensor, score: torch.Tensor,
               point: torch.Tensor, truth: Optional[torch.Tensor] = None, accuracy: Optional[torch.Tensor] = None) -> None:
    """Batched MultiTaskDecoder (+ _calculate_accuracy); see include/tpgan_b200.h."""
    B = count.shape[0]
    assert count.dtype == torch.int32 and count.numel() == B * num_classes
    assert score.numel() == B * num_classes * top_k and point.numel() == 2 * score.numel()
    _lib.check(_lib.load().tpgan_ssd_decode(_ptr(loc), _ptr(cls), B, n, loc_stride, cls_stride, num_classes, top_k,
                                            confidence_threshold, nms_distance, _ptr(count), _ptr(score), _ptr(point),
                                            _ptr(truth), _ptr(accuracy), _stream()), "ssd_decode")


# ---------------------------------------------------------------------------------------------------------- input pipeline
def u8_to_nhwc(src: torch.Tensor, dst: Act, round_tf32: bool = False) -> None:
    """uint8 (N,H,W,C) -> fp32 NHWC in [-1,1]: ToTensor()*2-1 (DataAndDataset.py:214-220)."""
    assert src.dtype == torch.uint8 and src.is_contiguous() and tuple(src.shape) == (dst.n, dst.h, dst.w, dst.c)
    _lib.check(_lib.load().tpgan_u8_to_nhwc(_ptr(src), dst.view(), int(round_tf32), _stream()), "u8_to_nhwc")


def landmarks_reduce(points: torch.Tensor, ranges: torch.Tensor, out: torch.Tensor, scale_x: float = 1.0,
                     scale_y: float = 1.0) -> None:
    """(N,P,2) landmark lists -> (N,R,2) means over inclusive index ranges (UtilityMethods.py:146-164)."""
    assert points.dtype == torch.float32 and points.is_contiguous() and ranges.dtype == torch.int32
    N, P = points.shape[0], points.shape[1]
    R = ranges.shape[0]
    assert tuple(out.shape) == (N, R, 2) and out.is_contiguous()
    _lib.check(_lib.load().tpgan_landmarks_reduce(_ptr(points), N, P, _ptr(ranges), R, scale_x, scale_y, _ptr(out), _stream()),
               "landmarks_reduce")


def pyramid(src: Act, half: Act, quarter: Act) -> None:
    """2x2 and 4x4 average pools of src in one pass."""
    _lib.check(_lib.load().tpgan_pyramid(src.view(), half.view(), quarter.view(), _stream()), "pyramid")
