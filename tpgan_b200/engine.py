"""Static-plan execution engine for the TP-GAN hot path.

The model code (tpgan_b200/D_and_G_model.py) is *traced once* per batch size: every layer call allocates its NHWC
activation buffer, records a forward kernel launch, and pushes a backward rule on a tape.  `Plan.trace_backward()` then
walks the tape in reverse and records the backward launches, deciding statically
  * whether a gradient contribution overwrites or accumulates (addend slots of the dgrad epilogue),
  * where the activation backward (LeakyReLU/ReLU mask) is fused: into the dgrad epilogue of the *last* contributor of
    each tensor (per-channel slopes for concat buffers), else one standalone mask kernel,
  * that residual-branch gradients ride as an extra addend of the sibling conv's dgrad (no separate add kernel).
Running a step is then just replaying two lists of prepared launches - no Python graph work, no allocation, no ATen
compute.  torch.cat never happens: producers write straight into channel slices of concat buffers.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops
from .ops import Act, CONV_DGRAD, CONV_FWD, DECONV_DGRAD, DECONV_FWD, EPI_LEAKY, EPI_LINEAR, EPI_MASK, round_up

LINEAR = 1.0  # "slope" of a tensor with no activation (nothing to mask in backward)


def dw_free(fn):
    """Mark a backward launch that neither reads nor writes packed weight gradients: the graph capture may leave the
    weight-gradient branch (train_step.GraphRunner, side stream) running across it."""
    fn.dw_free = True
    return fn


def _tag(fn, kind: str, nbytes: float, label: str = ""):
    """Attach (kind, algorithmic HBM bytes, label) to a recorded launch; bench.py groups per-launch timings by kind."""
    fn.kind, fn.bytes, fn.label = kind, nbytes, label
    return fn


class T:
    """A traced activation: forward view + static gradient bookkeeping."""

    def __init__(self, act: Act, slope: float = LINEAR, name: str = "", requires_grad: bool = True):
        self.act = act
        self.slope = slope
        self.name = name
        self.requires_grad = requires_grad
        self.parts: Optional[List["T"]] = None   # composite (concat buffer) = ordered channel segments
        self.parent: Optional["T"] = None
        self.cmap: Optional[List[int]] = None    # internal channel -> reference channel (-1 = padding lane)
        self.ref_c: int = act.c                  # number of reference (logical) channels
        self.grad: Optional[Act] = None
        self.flat = False                        # (B,1,1,C) tensor that the module API exposes as (B, C)
        # backward-trace state
        self.pending = 0
        self.grad_written = False
        self.masked = False
        self.deferred: List[Act] = []            # residual-branch addends waiting for a conv contribution
        # bf16 operand mode: is the bf16 twin of the activation / of its gradient up to date?  (tensor-core epilogues write
        # both copies; any other writer leaves the twin stale and the next tensor-core consumer inserts a cast)
        self.s16 = False
        self.g16 = False

    def leaves(self) -> List["T"]:
        return self.parts if self.parts is not None else [self]

    @property
    def shape(self):
        return (self.act.n, self.act.h, self.act.w, self.act.c)


class GradArena:
    """Bump allocator for the packed weight-gradient / bias-gradient accumulators of all layers on one device, so a step
    clears them with one memset per 256 MB chunk instead of two launches per layer."""

    _per_device: Dict[str, "GradArena"] = {}

    def __init__(self, device, chunk_elems: int = 1 << 26):
        self.device, self.chunk_elems = device, chunk_elems
        self.chunks: List[torch.Tensor] = []
        self.used: List[int] = []

    @classmethod
    def get(cls, device) -> "GradArena":
        key = str(torch.device(device))
        if key not in cls._per_device:
            cls._per_device[key] = GradArena(torch.device(device))
        return cls._per_device[key]

    def alloc(self, shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        n4 = round_up(max(n, 1), 64)  # 256-byte granularity keeps every allocation TMA/vector aligned
        if not self.chunks or self.used[-1] + n4 > self.chunks[-1].numel():
            self.chunks.append(torch.zeros(max(self.chunk_elems, n4), dtype=torch.float32, device=self.device))
            self.used.append(0)
        o = self.used[-1]
        self.used[-1] = o + n4
        return self.chunks[-1][o:o + n].view(*shape)

    def zero(self):
        for c, u in zip(self.chunks, self.used):
            c[:u].zero_()


class ConvLayer:
    """One Conv2d / ConvTranspose2d / Linear of the reference model: reference-layout parameters + tensor-core packings.

    `w_shape` lets a Linear be run as a convolution (fc1 = 8x8 valid conv over the 8x8x512 map, feature_predict.fc = 1x1)."""

    def __init__(self, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter], transposed: bool, k: int,
                 stride: int, pad: int, name: str = "", w_shape: Optional[Sequence[int]] = None):
        self.weight, self.bias = weight, bias
        self.w_shape = tuple(weight.shape) if w_shape is None else tuple(w_shape)
        self.transposed = transposed
        self.k, self.stride, self.pad = k, stride, pad
        self.name = name
        self.kind_fwd = DECONV_FWD if transposed else CONV_FWD
        self.kind_dgrad = DECONV_DGRAD if transposed else CONV_DGRAD
        self.cin = self.w_shape[0] if transposed else self.w_shape[1]
        self.cout = self.w_shape[1] if transposed else self.w_shape[0]
        self.bias_view = None
        self.ready = False

    def setup(self, in_cmap: Optional[List[int]], out_cmap: Optional[List[int]], in_c: int, out_c: int, device,
              need_dgrad: bool = True, exact: bool = False, defer_pack: bool = False, bf16: bool = False):
        """Allocate packings for the internal channel layouts seen at trace time (idempotent; layouts must not change).
        defer_pack: the caller re-packs all its layers with one multi-tensor launch (LayerSet.repack) right after tracing,
        so the per-layer pack / transpose launches here would be redundant."""
        key = (tuple(in_cmap) if in_cmap else None, tuple(out_cmap) if out_cmap else None, in_c, out_c)
        if self.ready:
            assert key == self._key, f"{self.name}: channel layout changed between traces"
            assert bf16 == self.bf16, f"{self.name}: a layer serves either tf32 or bf16 plans"
            if exact and self.wf_lo is None:
                self._alloc_lo()
                self.repack()
            return
        assert not (bf16 and exact), "the fp32-exact split mode is a tf32 verification mode"
        self.bf16 = bf16
        self.wf16 = self.wd16 = None
        self.wf_lo = self.wd_lo = None
        self._want_lo = exact
        self._key = key
        mk = lambda m: None if m is None else torch.tensor(m, dtype=torch.int32, device=device)
        self.in_map, self.out_map = mk(in_cmap), mk(out_cmap)
        self.in_c, self.out_c = in_c, out_c  # internal widths (incl. padding lanes when a map is given)
        self._alloc(device, need_dgrad)
        if bf16:    # bf16 operand copies of the packings (fp32 packings stay: they are the transpose / cast source)
            self.wf16 = ops.alloc_packed16(self.wf)
            self.wd16 = ops.alloc_packed16(self.wd) if self.wd is not None else None
        if exact:
            self._alloc_lo()
        nb = round_up(self.bias_len(), 4)
        self._bias_buf = ops.zeros(2 * nb, torch.float32, device) if self.bias is not None else None
        self.bias_int = self._bias_buf[:nb] if self.bias is not None else None
        self.db_int = GradArena.get(device).alloc((nb,)) if self.bias is not None else None
        if out_cmap is not None:
            oc = torch.tensor(out_cmap, dtype=torch.long, device=device)
            self._b_valid = (oc >= 0).nonzero().flatten()
            self._b_ref = oc[self._b_valid]
        self.ready = True
        if not defer_pack:
            self.repack()

    def bias_len(self):
        return self.out_c

    def _alloc(self, device, need_dgrad):
        shp = self.w_shape
        self.wf = ops.alloc_packed(self.kind_fwd, shp, rows_int=self.out_c, k_int=self.in_c, device=device)
        self.wd = ops.alloc_packed(self.kind_dgrad, shp, rows_int=self.in_c, k_int=self.out_c, device=device) \
            if need_dgrad else None
        self.dw = ops.Packed(GradArena.get(device).alloc(tuple(self.wf.data.shape)), self.wf.taps, self.wf.rows, self.wf.k,
                             self.wf.rows_pad, self.wf.k_pad)

    def _alloc_lo(self):
        """Residual packings w - tf32(w) for the fp32-exact verification mode (Plan(exact=True))."""
        cp = lambda pk: None if pk is None else ops.Packed(ops.zeros(pk.data.shape, torch.float32, pk.data.device), pk.taps, pk.rows, pk.k, pk.rows_pad,
                                                          pk.k_pad)
        self.wf_lo, self.wd_lo = cp(self.wf), cp(self.wd)

    def _w(self):
        return self.weight.data.view(self.w_shape)

    def repack(self):
        """Reference-layout fp32 master weights -> K-major packings (after every optimizer step): tf32-rounded fp32 for the
        tf32 kernels; un-rounded fp32 + their bf16 copies for the bf16 kernels."""
        if not self.ready:
            return
        self._repack32()
        if self.bf16:
            ops.cast_packed(self.wf, self.wf16)
            if self.wd is not None:
                ops.cast_packed(self.wd, self.wd16)

    def cast_jobs(self):
        """Multi-tensor form of the bf16 copies (LayerSet)."""
        if not self.bf16:
            return []
        return [ops.cast_job(self.wf, self.wf16)] + ([ops.cast_job(self.wd, self.wd16)] if self.wd is not None else [])

    def _repack32(self):
        w = self._w()
        rt = not self.bf16      # bf16 mode: keep fp32 bits here, round once when casting
        # the packing whose rows follow the reference's leading dimension is a row-contiguous (fast) pack; the other one
        # is its per-tap transpose
        if not self.transposed:
            ops.pack_weights(w, self.kind_fwd, self.wf, row_map=self.out_map, k_map=self.in_map, round_tf32=rt)
            if self.wd is not None:
                ops.transpose_packed(self.wf, self.wd)
        elif self.wd is not None:
            ops.pack_weights(w, self.kind_dgrad, self.wd, row_map=self.in_map, k_map=self.out_map, round_tf32=rt)
            ops.transpose_packed(self.wd, self.wf)
        else:
            ops.pack_weights(w, self.kind_fwd, self.wf, row_map=self.out_map, k_map=self.in_map, round_tf32=rt)
        if self.wf_lo is not None:
            ops.pack_weights(w, self.kind_fwd, self.wf_lo, row_map=self.out_map, k_map=self.in_map, round_tf32=2)
            if self.wd_lo is not None:
                ops.pack_weights(w, self.kind_dgrad, self.wd_lo, row_map=self.in_map, k_map=self.out_map, round_tf32=2)
        self._repack_bias()

    def _repack_bias(self):
        if self.bias is not None:
            if self.out_map is None:
                self.bias_int[: self.cout].copy_(self.bias.data)
            else:
                self.bias_int.zero_()
                self.bias_int[self._b_valid] = self.bias.data[self._b_ref]

    # ---- job descriptors for the multi-tensor launches (None = this layer needs its own calls)
    MAX_MULTI_ROW = 12288   # floats of shared memory per block in the multi-tensor pack kernel

    def _row_geometry(self):
        """(pack target, transpose target, rows map, k map, reference row length) of the row-contiguous packing."""
        taps = self.k * self.k
        if not self.transposed:
            return self.wf, self.wd, self.out_map, self.in_map, self.w_shape[1] * taps
        return self.wd, self.wf, self.in_map, self.out_map, self.w_shape[1] * taps

    def multi_ok(self) -> bool:
        if type(self) is not ConvLayer or not self.ready:
            return False
        if self.transposed and self.wd is None:
            return False
        return self._row_geometry()[4] <= self.MAX_MULTI_ROW

    def pack_jobs(self):
        """([pack jobs], [transpose jobs]) re-creating wf / wd / bias_int from the reference-layout parameters."""
        tgt, other, rmap, kmap, row_len = self._row_geometry()
        packs = [ops.pack_job(self.weight.data, tgt, self.k * self.k, row_len, rmap, kmap, 0 if self.bf16 else 1)]
        if self.bias is not None:
            nb = self.bias_int.numel()
            bp = ops.Packed(self._bias_buf, 1, 1, self.out_c, 1, nb)
            packs.append(ops.pack_job(self.bias.data, bp, 1, self.bias.numel(), None, self.out_map, 0))
        trans = [ops.transpose_job(tgt, other)] if other is not None else []
        if self.wf_lo is not None:   # fp32-exact mode: the residual packings w - tf32(w), same geometry (flag 2)
            tgt_lo, other_lo = (self.wf_lo, self.wd_lo) if not self.transposed else (self.wd_lo, self.wf_lo)
            packs.append(ops.pack_job(self.weight.data, tgt_lo, self.k * self.k, row_len, rmap, kmap, 2))
            if other_lo is not None:
                trans.append(ops.transpose_job(tgt_lo, other_lo))
        return packs, trans

    def unpack_jobs(self, accumulate: bool = False):
        """([transpose jobs], [unpack jobs]) writing weight.grad / bias.grad from the packed accumulators."""
        tgt, other, rmap, kmap, row_len = self._row_geometry()
        trans = []
        src = self.dw
        if self.transposed:
            if getattr(self, "_dw_t", None) is None:
                self._dw_t = ops.Packed(ops.zeros(self.wd.data.shape, torch.float32, self.wd.data.device), self.wd.taps, self.wd.rows, self.wd.k,
                                        self.wd.rows_pad, self.wd.k_pad)
            trans.append(ops.transpose_job(self.dw, self._dw_t))
            src = self._dw_t
        unp = [ops.pack_job(self.weight.grad, src, self.k * self.k, row_len, rmap, kmap, int(accumulate))]
        if self.bias is not None:
            nb = self.bias_int.numel()
            bp = ops.Packed(self.db_int, 1, 1, self.out_c, 1, nb)
            unp.append(ops.pack_job(self.bias.grad, bp, 1, self.bias.numel(), None, self.out_map, int(accumulate)))
        return trans, unp

    def _versions(self):
        return (self.weight._version, self.weight.data_ptr(), None if self.bias is None else self.bias._version)

    def refresh(self):
        """Repack if the reference-layout parameters changed since the last pack (optimizer.step, load_state_dict)."""
        if self.ready and getattr(self, "_packed_ver", None) != self._versions():
            self.repack()
            self._packed_ver = self._versions()

    def export_grad_autograd(self):
        """Accumulate this layer's gradients into param.grad (autograd semantics), creating it when absent."""
        for p in (self.weight, self.bias):
            if p is not None and p.requires_grad and p.grad is None:
                p.grad = torch.zeros_like(p)
        if self.weight.requires_grad:
            self.export_grad(accumulate=True)

    def zero_grad(self):
        if self.ready:
            self.dw.data.zero_()
            if self.db_int is not None:
                self.db_int.zero_()

    def export_grad(self, accumulate: bool = False):
        """Packed dW / internal db -> .grad of the reference-layout parameters."""
        if not self.ready or self.weight.grad is None:
            return
        if self.transposed and self.wd is not None:
            # deconv: transpose the [tap][Cout][Cin] gradient into the row-contiguous [tap][Cin][Cout] geometry first
            if getattr(self, "_dw_t", None) is None:
                self._dw_t = ops.Packed(ops.zeros(self.wd.data.shape, torch.float32, self.wd.data.device), self.wd.taps, self.wd.rows, self.wd.k,
                                        self.wd.rows_pad, self.wd.k_pad)
            ops.transpose_packed(self.dw, self._dw_t)
            ops.unpack_weights(self._dw_t, self.weight.grad.view(self.w_shape), self.kind_dgrad, row_map=self.in_map,
                               k_map=self.out_map, accumulate=accumulate)
        else:
            ops.unpack_weights(self.dw, self.weight.grad.view(self.w_shape), self.kind_fwd, row_map=self.out_map,
                               k_map=self.in_map, accumulate=accumulate)
        self._export_bias(accumulate)

    def _export_bias(self, accumulate):
        if self.bias is not None and self.bias.grad is not None:
            if self.out_map is None:
                src = self.db_int[: self.cout]
            else:
                src = torch.zeros_like(self.bias.grad)
                src[self._b_ref] = self.db_int[self._b_valid]
            if accumulate:
                self.bias.grad.add_(src)
            else:
                self.bias.grad.copy_(src)


class DeconvAsLinear(ConvLayer):
    """ConvTranspose2d(k, stride 1, pad 0) applied to a 1x1 map (deconv_8, D_and_G_model.py:218): a plain GEMM
    (B, Cin) x (Cin, k*k*Cout) whose NHWC output (B, k, k, Cout) is the flat row (tap, co).  Runs as a 1x1 convolution
    with N = k*k*Cout; the DECONV_FWD packing [tap][co][ci] *is* the [n = tap*Cout + co][ci] operand."""

    def __init__(self, weight, bias, k: int, name: str = ""):
        cin, cout = weight.shape[0], weight.shape[1]
        super().__init__(weight, bias, False, 1, 1, 0, name, w_shape=(k * k * cout, cin, 1, 1))
        self.ref_shape = tuple(weight.shape)
        self.kk, self.co = k * k, cout
        assert cout % 16 == 0
        self.bias_view = (k, k, cout)

    def bias_len(self):
        return self.co

    def _alloc(self, device, need_dgrad):
        assert self.out_map is None and self.in_map is None and self.in_c % 32 == 0
        n, kdim = self.kk * self.co, self.in_c
        mk = lambda: ops.zeros((2, n, kdim), torch.float32, device)
        self.wf = ops.Packed(mk(), 1, n, kdim, n, kdim)
        self.dw = ops.Packed(GradArena.get(device).alloc((2, n, kdim)), 1, n, kdim, n, kdim)
        # views of the same storage in the [taps+1][Cout][Cin] geometry the (un)pack kernels use
        self._wf_taps = ops.Packed(self.wf.data, self.kk, self.co, kdim, self.co, kdim)
        self._dw_taps = ops.Packed(self.dw.data, self.kk, self.co, kdim, self.co, kdim)
        self.wd = ops.Packed(ops.zeros((2, round_up(kdim, 16), n), torch.float32, device), 1, kdim, n,
                             round_up(kdim, 16), n)
        kk, co = self.kk, self.co
        idx = torch.arange(n, dtype=torch.int64)
        self._kmap_d = ((idx % co) * kk + idx // co).to(torch.int32).to(device)  # (tap, co) -> co*kk + tap
        self.bias_full = ops.zeros(n, torch.float32, device)

    def _alloc_lo(self):
        super()._alloc_lo()
        self._wf_lo_taps = ops.Packed(self.wf_lo.data, self.kk, self.co, self.in_c, self.co, self.in_c)

    def _repack32(self):
        w = self.weight.data
        rt = 0 if self.bf16 else 1
        ops.pack_weights(w, DECONV_FWD, self._wf_taps, round_tf32=rt)
        lib = ops._lib.load()
        n = self.kk * self.co
        ops._lib.check(lib.tpgan_pack_weights(w.data_ptr(), self.wd.data.data_ptr(), 1, self.in_c, n, self.wd.rows_pad, n,
                                              n, 1, None, self._kmap_d.data_ptr(), rt, ops._stream()), "pack deconv_8 dgrad")
        if self.wf_lo is not None:
            ops.pack_weights(w, DECONV_FWD, self._wf_lo_taps, round_tf32=2)
            ops._lib.check(lib.tpgan_pack_weights(w.data_ptr(), self.wd_lo.data.data_ptr(), 1, self.in_c, n,
                                                  self.wd.rows_pad, n, n, 1, None, self._kmap_d.data_ptr(), 2,
                                                  ops._stream()), "pack deconv_8 dgrad lo")
        if self.bias is not None:
            self.bias_full.view(self.kk, self.co).copy_(self.bias.data.view(1, self.co).expand(self.kk, self.co))

    def export_grad(self, accumulate: bool = False):
        if not self.ready or self.weight.grad is None:
            return
        ops.unpack_weights(self._dw_taps, self.weight.grad, DECONV_FWD, accumulate=accumulate)
        if self.bias is not None and self.bias.grad is not None:
            src = self.db_int[: self.co]
            self.bias.grad.add_(src) if accumulate else self.bias.grad.copy_(src)


class Plan:
    """Traced schedule for one network instance and batch size."""

    def __init__(self, device, training: bool = True, need_wgrad: bool = True, exact: bool = False,
                 defer_bias: bool = False, defer_pack: bool = False, bf16: bool = False):
        self.device = device
        # bf16 operand mode (BASELINE configs[2]): tensor-core operands are the bf16 twins of the activations / activation
        # gradients (ops.Arena(shadow=True)), written by the conv epilogues next to the fp32 copies that the pointwise
        # kernels, the masks and the residual / gradient accumulations keep using; fp32 accumulate, fp32 master weights
        self.bf16 = bf16
        assert not (bf16 and exact)
        self.twin_reads: List[Act] = []   # every view a tensor-core launch of this plan reads through its bf16 twin (tests:
                                          # after a step each twin must equal the bf16 rounding of its fp32 copy)
        self.defer_pack = defer_pack  # the owner re-packs every layer of this plan in one multi-tensor launch after tracing
        self.defer_bias = defer_bias  # record (layer, dY, ready index) instead of launching one bias-grad kernel per layer
        self.bias_jobs: list = []
        self.exact = exact  # fp32-exact verification mode: every tensor-core product is split hi/lo (3 launches)
        self.training = training
        self.need_wgrad = need_wgrad
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self.tape: List[Callable[[], None]] = []
        self.grad_bufs: Dict[int, torch.Tensor] = {}
        self.keep: list = []  # keeps buffers / ctypes structs alive
        self.bytes = 0
        self.bwd_marks: Dict[str, int] = {}   # layer name -> index in self.bwd after which its dW is final
        self.layers: List[ConvLayer] = []     # every conv layer this plan launches (for repack / grad export)
        self.named: Dict[str, T] = {}         # layer name -> its output tensor (introspection / tests)
        self.aux: list = []                   # non-tensor-core parameter holders (BatchNorm, depthwise conv) of this plan
        self.direct_grads = False             # aux layers write straight into param.grad (flat buffers of a trainer)
        self.bn_training: Optional[bool] = None   # BatchNorm mode when it differs from `training` (module.train() under no_grad)

    # ------------------------------------------------------------------ buffers
    def new(self, n, h, w, c, slope=LINEAR, name="", requires_grad=True) -> T:
        a = Act.empty(n, h, w, c, self.device)
        self.bytes += a.buf.numel() * 4
        return T(a, slope, name, requires_grad)

    def wrap(self, act: Act, slope=LINEAR, name="", requires_grad=True, s16: bool = False) -> T:
        """s16: the bf16 twin of `act` is already current (it was written by a tensor-core epilogue of another plan)."""
        t = T(act, slope, name, requires_grad)
        t.s16 = s16
        return t

    def concat(self, n, h, w, widths: Sequence[int], name="") -> T:
        """Concat buffer: parts are written in place by their producers (replaces torch.cat, D_and_G_model.py:100...)."""
        offs, o = [], 0
        for c in widths:
            offs.append(o)
            o += round_up(c, 4)
        buf = ops.zeros((n, h, w, round_up(o, 8) if self.bf16 else o), torch.float32, self.device)
        self.bytes += buf.numel() * 4
        padded = any(c % 4 for c in widths[:-1])
        total_c = o if padded else offs[-1] + widths[-1]
        whole = T(Act(buf, 0, total_c), LINEAR, name)
        whole.parts = []
        cmap, ref = [], 0
        for c, off in zip(widths, offs):
            p = T(Act(buf, off, c), LINEAR, f"{name}[{off}:{off + c}]")
            p.parent = whole
            whole.parts.append(p)
            cmap += list(range(ref, ref + c)) + [-1] * (round_up(c, 4) - c)
            ref += c
        whole.ref_c = ref
        whole.cmap = cmap[:total_c] if padded else None
        return whole

    def grad_act(self, t: T) -> Act:
        if t.grad is None:
            buf = t.act.buf
            g = self.grad_bufs.get(id(buf))
            if g is None:
                g = ops.zeros(buf.shape, torch.float32, buf.device)
                self.grad_bufs[id(buf)] = g
                self.bytes += g.numel() * 4
            t.grad = Act(g, t.act.c0, t.act.c)
        return t.grad

    # ------------------------------------------------------------------ forward ops
    def use(self, t: T):
        """Register one differentiable consumer of t (one future gradient contribution)."""
        if t.requires_grad:
            for p in t.leaves():
                p.pending += 1

    def conv(self, layers: Sequence[ConvLayer], xs: Sequence[T], slope: Optional[float], outs: Optional[Sequence[T]] = None,
             residuals: Optional[Sequence[Optional[T]]] = None, name: str = "", round_out: bool = True) -> List[T]:
        """Grouped conv/deconv forward (+bias, +residual, +activation).  slope None = no activation.  round_out=False
        keeps the stored result in full fp32 (consumers that are not tensor-core operands: BatchNorm, losses)."""
        G = len(layers)
        res = list(residuals) if residuals is not None else [None] * G
        outs_l: List[T] = []
        args = []
        for i, (L, x) in enumerate(zip(layers, xs)):
            n, h, w = x.act.n, x.act.h, x.act.w
            if L.transposed:
                ho = (h - 1) * L.stride - 2 * L.pad + L.k + (1 if L.stride > 1 else 0)
                wo = (w - 1) * L.stride - 2 * L.pad + L.k + (1 if L.stride > 1 else 0)
            else:
                ho, wo = (h + 2 * L.pad - L.k) // L.stride + 1, (w + 2 * L.pad - L.k) // L.stride + 1
            r = res[i]
            if outs is not None and outs[i] is not None:
                out = outs[i]
                assert (out.act.n, out.act.h, out.act.w) == (n, ho, wo), (L.name, out.shape, (n, ho, wo))
            else:
                oc = r.act.c if r is not None else L.cout
                out = self.new(n, ho, wo, oc, name=L.name)
            out_cmap = None
            if r is not None:  # residual output lives in the input tensor's internal channel layout
                assert r.act.c == out.act.c
                out_cmap = r.cmap
                out.cmap, out.ref_c = r.cmap, r.ref_c
            out.slope = LINEAR if slope is None else slope
            L.setup(x.cmap, out_cmap, x.act.c, out.act.c, self.device, exact=self.exact, defer_pack=self.defer_pack,
                    bf16=self.bf16)
            if L not in self.layers:
                self.layers.append(L)
            self.named[L.name] = out
            assert (x.ref_c == L.cin) and (out.ref_c == L.cout), (L.name, x.ref_c, L.cin, out.ref_c, L.cout)
            args.append(dict(L=L, dgrad=False, x=x.act, out=out.act, bias=getattr(L, "bias_full", L.bias_int),
                             add1=None if r is None else r.act, slope=0.0 if slope is None else slope,
                             epilogue=EPI_LINEAR if slope is None else EPI_LEAKY, round=round_out))
            self.use(x)
            if r is not None:
                self.use(r)
            outs_l.append(out)
            self._ensure16(x, self.fwd)
        self._emit_conv(args, self.fwd)
        for out in outs_l:
            for p in out.leaves():
                p.s16 = self.bf16
        self.tape.append(lambda: self._bwd_conv(layers, list(xs), outs_l, res))
        return outs_l

    # ------------------------------------------------------------------ bf16 twins
    def _ensure16(self, t: T, lst):
        """Make the bf16 twin of activation t current before a tensor-core consumer reads it."""
        if not self.bf16:
            return
        for p in t.leaves():
            self.twin_reads.append(p.act)
            if not p.s16:
                a = p.act
                lst.append(_tag(lambda a=a: ops.cast_bf16(a), "cast16", 6.0 * a.n * a.h * a.w * a.c, p.name))
                p.s16 = True

    def _ensure_g16(self, t: T):
        """The same for the gradient of t (read by dgrad and wgrad launches as their dY operand)."""
        if not self.bf16:
            return
        for p in t.leaves():
            self.twin_reads.append(self.grad_act(p))
            if not p.g16:
                g = self.grad_act(p)
                self.bwd.append(_tag(lambda g=g: ops.cast_bf16(g), "cast16", 6.0 * g.n * g.h * g.w * g.c, p.name + ".grad"))
                p.g16 = True

    @staticmethod
    def _flops(L, x: Act, out: Act, dgrad: bool) -> float:
        """Algorithmic FLOPs (2*MACs, unpadded reference channel counts) of one conv-like launch."""
        pix_src = (out if L.transposed else x) if dgrad else (x if L.transposed else out)
        return 2.0 * pix_src.n * pix_src.h * pix_src.w * L.cin * L.cout * L.k * L.k

    def _split(self, x: Act, lst):
        """fp32-exact mode: tf32 high part and residual of `x`, emitted into `lst`.  The input-gradient and weight-gradient
        launches of a layer are emitted back to back and both read the same (final) output gradient: the second request
        for the same tensor in the same list reuses the first split instead of launching it again."""
        key = (x.buf.data_ptr(), x.c0, x.c, tuple(x.buf.shape), id(lst))
        recent = getattr(self, "_recent_splits", [])
        for k, pos, xh, xl in recent:
            if k == key and pos >= len(lst) - 4:     # emitted a launch or two ago: nothing in between writes the tensor
                return xh, xl
        # a layer input split for the forward launch is still valid when its weight gradient is traced: forward
        # activations are never written again within a step (the weight gradient itself relies on that)
        fwd_splits = self.__dict__.setdefault("_fwd_splits", {})
        if lst is self.bwd and key[:4] in fwd_splits:
            return fwd_splits[key[:4]]
        xh, xl = Act.empty(x.n, x.h, x.w, x.c, self.device), Act.empty(x.n, x.h, x.w, x.c, self.device)
        self.keep.append((xh, xl))
        lst.append(dw_free(lambda x=x, xh=xh, xl=xl: ops.split_tf32(x, xh, xl)))
        self._recent_splits = (recent + [(key, len(lst), xh, xl)])[-2:]
        if lst is self.fwd:
            fwd_splits[key[:4]] = (xh, xl)
        return xh, xl

    def _emit_conv(self, specs, lst):
        """specs: dicts {L, dgrad, x, out, bias, add1, add2, mask, slopes, slope, epilogue} of one grouped launch."""
        fl = sum(self._flops(sp["L"], sp["x"], sp["out"], sp["dgrad"]) for sp in specs)
        def mk(sp, x, out, pack, x_lo=None, w_lo=None, **kw):
            L = sp["L"]
            return ops.conv_args(L.kind_dgrad if sp["dgrad"] else L.kind_fwd, x, out, pack, L.k, L.stride, L.pad,
                                 round_tf32=(not self.exact) and sp.get("round", True), bf16=self.bf16,
                                 out16=sp.get("out16", True), x_lo=x_lo, w_lo=w_lo, **kw)
        full = lambda sp: dict(bias=sp.get("bias"), add1=sp.get("add1"), add2=sp.get("add2"), mask=sp.get("mask"),
                               slopes=sp.get("slopes"), slope=sp.get("slope", 0.0), epilogue=sp.get("epilogue", EPI_LINEAR))
        if not self.exact:
            # tf32: fp32 activations + tf32-rounded packings; bf16: the bf16 twins of the activations + bf16 packings
            wpack = (lambda sp: sp["L"].wd16 if sp["dgrad"] else sp["L"].wf16) if self.bf16 else \
                (lambda sp: sp["L"].wd if sp["dgrad"] else sp["L"].wf)
            run = self._conv_launch([mk(sp, sp["x"], sp["out"], wpack(sp), **full(sp)) for sp in specs], fl,
                                    ",".join(sp["L"].name for sp in specs) + (":dgrad" if specs[0]["dgrad"] else ":fwd"))
            sp0, L0 = specs[0], specs[0]["L"]
            pk = L0.wd if sp0["dgrad"] else L0.wf
            if (len(specs) == 1 and not L0.transposed and L0.stride == 1 and 2 * L0.pad == L0.k - 1 and L0.k >= 3
                    and sp0["x"].w == 128 and sp0["out"].w == 128 and pk.rows_pad <= 256):
                run.kind = "rowconv"   # the library routes these to rowconv_kernel (api.cu: try_rowconv)
            lst.append(run)
            return
        # fp32-exact mode.  Kernels up to 4x4 (every MobileNetV2 layer): the three-term product a_h*w_l + a_l*w_h + a_h*w_h
        # is ONE launch - the tap list of the launch three times over the hi / lo operand copies (tpgan_conv_args.in_lo) -
        # after one launch that splits the activation.  Larger kernels (3 x k*k taps exceed the tap table): three chained
        # launches through a scratch tensor, as before.
        def fits(L):   # taps of the launch (hole phases of a stride > kernel deconv add one zero tap each), three times
            holes = L.stride ** 2 if (L.transposed and L.stride > L.k) else 0
            return 3 * (L.k ** 2 + holes) <= 64
        if all(fits(sp["L"]) for sp in specs):
            args = []
            for sp in specs:
                L, x, out = sp["L"], sp["x"], sp["out"]
                hi_pack, lo_pack = (L.wd, L.wd_lo) if sp["dgrad"] else (L.wf, L.wf_lo)
                xh, xl = self._split(x, lst)
                args.append(mk(sp, xh, out, hi_pack, x_lo=xl, w_lo=lo_pack, **full(sp)))
            run = self._conv_launch(args, fl, ",".join(sp["L"].name for sp in specs) + (":dgrad" if specs[0]["dgrad"] else ":fwd"))
            lst.append(run)
            return
        a1, a2, a3 = [], [], []
        for sp in specs:
            L, x, out = sp["L"], sp["x"], sp["out"]
            hi_pack, lo_pack = (L.wd, L.wd_lo) if sp["dgrad"] else (L.wf, L.wf_lo)
            xh, xl = Act.empty(x.n, x.h, x.w, x.c, self.device), Act.empty(x.n, x.h, x.w, x.c, self.device)
            tmp = Act.empty(out.n, out.h, out.w, out.c, self.device)
            self.keep.append((xh, xl, tmp))  # ConvArgs hold raw pointers only
            lst.append(lambda x=x, xh=xh, xl=xl: ops.split_tf32(x, xh, xl))
            a1.append(mk(sp, xh, tmp, lo_pack))
            a2.append(mk(sp, xl, tmp, hi_pack, add1=tmp))
            kw = full(sp)
            if kw["add1"] is not None and kw["add2"] is not None:
                extra = kw["add2"]
                a2_post = (lambda e=extra, t=tmp: ops.view_copy(e, t, True))
                kw["add2"] = tmp
                a3.append((mk(sp, xh, out, hi_pack, **kw), a2_post))
            else:
                if kw["add1"] is None:
                    kw["add1"] = tmp
                else:
                    kw["add2"] = tmp
                a3.append((mk(sp, xh, out, hi_pack, **kw), None))
        lst.append(self._conv_launch(a1))
        lst.append(self._conv_launch(a2))
        for _, post in a3:
            if post is not None:
                lst.append(post)
        lst.append(self._conv_launch([a for a, _ in a3]))

    def _emit_wgrad(self, specs, lst, accumulate: bool = True):
        """specs: (L, x Act, dy Act) of one grouped weight-gradient launch.  accumulate=False: the launch is the only
        writer of each layer's (cleared) dW this step."""
        acc = accumulate or self.exact
        mk = lambda L, x, dy: ops.wgrad_args(L.kind_fwd, x, dy, L.dw, L.k, L.stride, L.pad, acc, bf16=self.bf16)
        if not self.exact:
            fl = sum(self._flops(L, x, dy, False) for L, x, dy in specs)
            lst.append(self._wgrad_launch([mk(L, x, dy) for L, x, dy in specs], fl, ",".join(L.name for L, _, _ in specs) + ":wgrad"))
            return
        g1, g2, g3 = [], [], []
        for L, x, dy in specs:
            dh, dl = self._split(dy, lst)
            xh, xl = self._split(x, lst)
            g1.append(mk(L, xh, dh))
            g2.append(mk(L, xl, dh))
            g3.append(mk(L, xh, dl))
        if len(specs) == 1:
            # the three products of one layer accumulate into the same dW: ONE grouped launch of three problems (fp32
            # reds into the cleared dW; a grouped launch takes up to four independent problems)
            lst.append(self._wgrad_launch(g1 + g2 + g3, sum(self._flops(L, x, dy, False) for L, x, dy in specs),
                                          specs[0][0].name + ":wgrad"))
            return
        for g in (g1, g2, g3):
            lst.append(self._wgrad_launch(g))

    def _conv_launch(self, args, flops: float = 0.0, label: str = ""):
        arr = (ops.ConvArgs * len(args))(*args)
        self.keep.append(arr)
        n = len(args)
        lib = ops._lib.load()

        def run():
            ops._lib.check(lib.tpgan_conv2d(arr, n, ops._stream()), "conv2d")
        run.kind, run.flops, run.label = "tapgemm", flops, label
        return run

    def _wgrad_launch(self, args, flops: float = 0.0, label: str = ""):
        arr = (ops.WgradArgs * len(args))(*args)
        self.keep.append(arr)
        n = len(args)
        lib = ops._lib.load()

        def run():
            ops._lib.check(lib.tpgan_conv2d_wgrad(arr, n, ops._stream()), "wgrad")
        run.kind, run.flops, run.label = "wgrad", flops, label
        return run

    def reflect_pad(self, x: T, left: int, top: int) -> T:
        out = self.new(x.act.n, x.act.h + top, x.act.w + left, x.act.c, slope=LINEAR, name=x.name + ".rpad")
        out.cmap, out.ref_c = x.cmap, x.ref_c
        self.use(x)
        self.fwd.append(lambda: ops.reflect_pad(x.act, out.act, left, top))

        def bwd():
            if not self._has_grad(out):
                return self._null(x)
            g = self.grad_act(out)
            self._contribute(x, lambda dst, acc: ops.reflect_pad_backward(g, dst, left, top, acc))
        self.tape.append(bwd)
        return out

    def copy(self, src: T, dst: T):
        """dst <- src (view copy); gradient flows back dst -> src."""
        assert src.act.c == dst.act.c
        dst.slope = LINEAR
        self.use(src)
        self.fwd.append(lambda: ops.view_copy(src.act, dst.act, False))

        def bwd():
            if not self._has_grad(dst):
                return self._null(src)
            self._finalize(dst)
            g = self.grad_act(dst)
            self._contribute(src, lambda d, acc: ops.view_copy(g, d, acc))
        self.tape.append(bwd)

    def custom(self, fwd: Callable[[], None], inputs: Sequence[T], outputs: Sequence[T],
               bwd: Callable[[List[Optional[Callable]]], None]):
        """Generic op: `bwd(contribs)` is called at backward-trace time only if some output has a gradient; contribs[i]
        is a function emit(fn(dst_act, accumulate)) registering input i's gradient contribution."""
        for x in inputs:
            self.use(x)
        self.fwd.append(fwd)

        def tape_fn():
            if not any(self._has_grad(o) for o in outputs):
                for x in inputs:
                    self._null(x)
                return
            for o in outputs:
                if self._has_grad(o):
                    self._finalize(o)
            contribs = [(lambda fn, x=x: self._contribute(x, fn)) for x in inputs]
            bwd(contribs)
        self.tape.append(tape_fn)

    def alias(self, src: T, h: int, w: int, c: int, name: str = "") -> T:
        """Same storage seen with another (h, w, c) factorisation of the per-image row (e.g. deconv_8's (1,1,4096) GEMM row
        = the (8,8,64) NHWC map).  `src` must have no other consumer; the alias has no activation of its own."""
        buf = src.act.buf
        n = buf.shape[0]
        assert src.act.c0 == 0 and src.act.c == buf.shape[3] and buf.shape[1] * buf.shape[2] * buf.shape[3] == h * w * c
        nbuf = buf.view(n, h, w, c)
        out = T(Act(nbuf, 0, c), LINEAR, name or src.name + ".alias")
        out.s16 = src.s16           # same storage, same twin
        self.use(src)
        self.keep.append(nbuf)

        def bwd():
            if not self._has_grad(out):
                return self._null(src)
            self._finalize(out)
            assert not src.grad_written, "alias source must have a single consumer"
            src.grad = Act(self.grad_act(out).buf.view(buf.shape), 0, src.act.c)
            src.grad_written = True
            src.g16 = out.g16
            src.pending -= 1
        # the two T's must share one gradient buffer: allocate it through the alias and view it back
        self.tape.append(bwd)
        return out

    def local_fuse(self, parts: Sequence[T], out: T, name: str = "fuse") -> T:
        """LocalFuser (D_and_G_model.py:132-159): max over the four zero-padded patches at the fixed offsets."""
        need_grad = any(p.requires_grad for p in parts)
        n, c = out.act.n, out.act.c
        argmax = torch.empty((n, 128, 128, c), dtype=torch.uint8, device=self.device) if need_grad else None
        self.bytes += 0 if argmax is None else argmax.numel()
        out.slope = LINEAR
        acts = [p.act for p in parts]
        for p in parts:
            self.use(p)
        self.fwd.append(lambda: ops.local_fuse(acts, out.act, argmax))
        self.fuse_argmax = getattr(self, "fuse_argmax", {})
        self.fuse_argmax[name] = argmax

        def bwd():
            if not need_grad or not self._has_grad(out):
                for p in parts:
                    self._null(p)
                return
            self._finalize(out)
            g = self.grad_act(out)
            written = [all(l.grad_written for l in p.leaves()) for p in parts]
            acc = any(written)
            if acc and not all(written):
                for p in parts:
                    self._zero_unwritten(p)
            dsts = [self.grad_act(p) for p in parts]
            self.bwd.append(lambda: ops.local_fuse_backward(g, argmax, dsts, acc))
            for p in parts:
                for l in p.leaves():
                    l.grad_written = True
                    l.g16 = False
                    l.pending -= 1
        self.tape.append(bwd)
        return out

    def maxout2(self, x: T, name: str = "maxout") -> T:
        """nn.MaxPool1d(2,2) over adjacent feature pairs of a (B,1,1,2C) row (D_and_G_model.py:214,290)."""
        n, c2 = x.act.n, x.act.c
        assert x.act.h == 1 and x.act.w == 1 and c2 % 8 == 0 and x.act.c0 == 0 and x.act.buf.shape[3] == c2
        out = self.new(n, 1, 1, c2 // 2, name=name)
        self.use(x)
        xb, yb = x.act.buf.view(n, c2), out.act.buf.view(n, c2 // 2)
        self.fwd.append(lambda: ops.maxout2(xb, yb))

        def bwd():
            if not self._has_grad(out):
                return self._null(x)
            self._finalize(out)
            g = self.grad_act(out).buf.view(n, c2 // 2)

            def emit(dst, acc):
                assert not acc
                ops.maxout2_backward(xb, g, dst.buf.view(n, c2))
            self._contribute(x, emit)
        self.tape.append(bwd)
        return out

    def mul_mask(self, x: T, mask: Act, name: str = "dropout") -> T:
        """y = x * mask (nn.Dropout with an externally drawn, pre-scaled mask; D_and_G_model.py:344-346)."""
        out = self.new(x.act.n, x.act.h, x.act.w, x.act.c, name=name)
        self.use(x)
        self.fwd.append(lambda: ops.mul(x.act, mask, out.act))

        def bwd():
            if not self._has_grad(out):
                return self._null(x)
            self._finalize(out)
            g = self.grad_act(out)

            def emit(dst, acc):
                assert not acc
                ops.mul(g, mask, dst)
            self._contribute(x, emit)
        self.tape.append(bwd)
        return out

    def maxpool3s2(self, x: T, name: str = "maxpool") -> T:
        """nn.MaxPool2d(3, 2, 1) (identity network, ResNet.py:33)."""
        n, h, w, c = x.act.n, x.act.h, x.act.w, x.act.c
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        out = self.new(n, ho, wo, c, name=name)
        need = x.requires_grad and self.training
        arg = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=self.device) if need else None
        self.use(x)
        self.fwd.append(lambda: ops.maxpool3s2(x.act, out.act, arg))

        def bwd():
            if not need or not self._has_grad(out):
                return self._null(x)
            self._finalize(out)
            g = self.grad_act(out)
            self._contribute(x, lambda dst, acc: ops.maxpool3s2_backward(g, arg, dst, acc))
        self.tape.append(bwd)
        return out

    def avgpool(self, x: T, name: str = "avgpool") -> T:
        """nn.AdaptiveAvgPool2d((1, 1)) (ResNet.py:45)."""
        out = self.new(x.act.n, 1, 1, x.act.c, name=name)
        self.use(x)
        self.fwd.append(lambda: ops.avgpool(x.act, out.act))

        def bwd():
            if not self._has_grad(out):
                return self._null(x)
            self._finalize(out)
            g = self.grad_act(out)
            self._contribute(x, lambda dst, acc: ops.avgpool_backward(g, dst, acc))
        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ Pretrain path (MobileNetV2.py) ops
    def batchnorm(self, bn, x: T, res: Optional[T] = None, relu6: bool = False, round_out: bool = True,
                  round_dx: bool = True, name: str = "", slope: Optional[float] = None, out: Optional[T] = None) -> T:
        """nn.BatchNorm2d (+nn.ReLU6, + the residual add of InvertedResidual.forward, MobileNetV2.py:107-120).  `bn` is a
        BNLayer (tpgan_b200/MobileNetV2.py).  Training plans use batch statistics and update the running ones."""
        n, h, w, c = x.shape
        assert c % 4 == 0 and x.act.c0 == 0 and x.act.buf.shape[3] == c, "BatchNorm needs a whole pixel-dense buffer"
        if out is None:
            out = self.new(n, h, w, c, name=name or bn.name)
        else:   # write straight into a pre-allocated tensor (a channel slice of a concat buffer): pixel-dense, 16-byte lanes
            assert (out.act.n, out.act.h, out.act.w, out.act.c) == (n, h, w, c) and out.act.c0 % 4 == 0, (out.shape, x.shape)
            out.slope = LINEAR
        if slope is not None:      # (Leaky)ReLU after the BatchNorm (+ residual): forward in the kernel, backward through
            assert not relu6       # the engine's activation-mask machinery (sign of the stored output)
            out.slope = slope
        st = bn.state(self, c)
        if bn not in self.aux:
            self.aux.append(bn)
        training = self.training if self.bn_training is None else self.bn_training
        rt = round_out and not self.exact
        self.use(x)
        if res is not None:
            self.use(res)
        m = bn.module
        eps = float(m.eps)
        mom = 0.1 if m.momentum is None else float(m.momentum)
        elems = n * h * w * c
        self.fwd.append(_tag(lambda: ops.bn_forward(x.act, None if res is None else res.act, out.act, m.weight.data,
                                                    m.bias.data, m.running_mean, m.running_var, mom, eps, training, relu6,
                                                    rt, st.sums, st.coef, slope),
                             "bn_fwd", 4.0 * elems * ((3 if training else 2) + (res is not None)), name or bn.name))
        self.named[name or bn.name] = out

        def tape_fn():
            if not self._has_grad(out):
                self._null(x)
                if res is not None:
                    self._null(res)
                return
            self._finalize(out)
            g = self.grad_act(out)
            need_param = self.need_wgrad       # eval mode too: frozen statistics, trainable gamma / beta
            dg, db = (bn.grad_targets(self) if need_param else (None, None))
            rdx = round_dx and not self.exact
            if res is not None:
                self._contribute(res, lambda dst, acc: ops.view_copy(g, dst, acc))
            if x.requires_grad:
                self._contribute(x, _tag(lambda dst, acc: ops.bn_backward(g, x.act, dst, st.coef, training, relu6, acc, rdx,
                                                                          st.dsums, dg, db),
                                         "bn_bwd", 4.0 * elems * (5 if training else 3), name or bn.name))
        self.tape.append(tape_fn)
        return out

    def dwconv(self, layer, x: T, name: str = "") -> T:
        """Depthwise 3x3 conv, pad 1, stride 1|2 (MobileNetV2.py:110).  `layer` is a DepthwiseLayer."""
        n, h, w, c = x.shape
        s = layer.stride
        ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
        out = self.new(n, ho, wo, c, name=name or layer.name)
        if layer not in self.aux:
            self.aux.append(layer)
        wt = layer.weight
        self.use(x)
        ein, eout = n * h * w * c, n * ho * wo * c
        self.fwd.append(_tag(lambda: ops.dwconv3x3(x.act, out.act, wt.data, s), "dw_fwd", 4.0 * (ein + eout), layer.name))
        self.named[name or layer.name] = out

        def tape_fn():
            if not self._has_grad(out):
                return self._null(x)
            self._finalize(out)
            g = self.grad_act(out)
            if self.need_wgrad:
                dw = layer.grad_target(self)
                self.bwd.append(_tag(lambda: ops.dwconv3x3_wgrad(x.act, g, dw, s), "dw_wgrad", 4.0 * (ein + eout), layer.name))
            if x.requires_grad:
                self._contribute(x, _tag(lambda dst, acc: ops.dwconv3x3_dgrad(g, dst, wt.data, s, acc), "dw_dgrad",
                                         4.0 * (ein + eout), layer.name))
        self.tape.append(tape_fn)
        return out

    def gather_rows(self, parts: Sequence[T], name: str = "rows") -> T:
        """SSDHead.forward's view(N,-1,K) + torch.cat(dim=1) (MobileNetV2.py:62-76): the NHWC head outputs laid end to end
        per image.  Returns a flat (N,1,1,total) tensor."""
        n = parts[0].act.n
        sizes = [p.act.h * p.act.w * p.act.c for p in parts]
        total = sum(sizes)
        out = self.new(n, 1, 1, total, name=name)
        out.flat = True
        stride = out.act.buf.shape[3]
        offs = [sum(sizes[:i]) for i in range(len(sizes))]
        for p in parts:
            self.use(p)
        acts = [p.act for p in parts]
        self.fwd.append(lambda: [ops.rows_gather(a, out.act.buf, stride, o) for a, o in zip(acts, offs)])

        def tape_fn():
            if not self._has_grad(out):
                for p in parts:
                    self._null(p)
                return
            self._finalize(out)
            g = self.grad_act(out).buf
            for p, o in zip(parts, offs):
                def emit(dst, acc, o=o):
                    assert not acc, "a head output has a single consumer"
                    ops.rows_gather(dst, g, stride, o, reverse=True)
                self._contribute(p, emit)
        self.tape.append(tape_fn)
        return out

    # ------------------------------------------------------------------ backward tracing
    def seed_grad(self, t: T):
        """Declare that an external kernel (a loss) writes t's gradient before the backward replay."""
        self.grad_act(t)
        for p in t.leaves():
            p.grad_written = True
            p.g16 = False

    def trace_backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    def _has_grad(self, t: T) -> bool:
        return any(p.grad_written or p.deferred for p in t.leaves()) or bool(t.deferred)

    def _null(self, x: T):
        if x.requires_grad:
            for p in x.leaves():
                p.pending -= 1

    def _flush_deferred(self, t: T):
        """Residual addends that no conv contribution picked up: add them explicitly."""
        for holder in ([t] + ([t.parent] if t.parent is not None else [])):
            while holder.deferred:
                add = holder.deferred.pop()
                dst = self.grad_act(holder)
                acc = all(p.grad_written for p in holder.leaves())
                if not acc and any(p.grad_written for p in holder.leaves()):
                    self._zero_unwritten(holder)
                    acc = True
                self.bwd.append(lambda add=add, dst=dst, acc=acc: ops.view_copy(add, dst, acc))
                for p in holder.leaves():
                    p.grad_written = True
                    p.g16 = False

    def _zero_unwritten(self, t: T):
        for p in t.leaves():
            if not p.grad_written:
                g = self.grad_act(p)
                self.bwd.append(dw_free(lambda g=g: g.buf[..., g.c0:g.c0 + g.c].zero_()))
                p.grad_written = True
                p.g16 = False

    def _finalize(self, t: T):
        """Make t.grad hold d(loss)/d(pre-activation): flush deferred addends, apply the activation mask if nobody fused
        it."""
        self._flush_deferred(t)
        for p in t.leaves():
            if p.slope != LINEAR and not p.masked and p.grad_written:
                g = self.grad_act(p)
                self.bwd.append(dw_free(lambda g=g, p=p: ops.act_backward(g, p.act, g, slope=p.slope)))
                p.masked = True
                p.g16 = False

    def _contribute(self, x: T, emit: Callable[[Act, bool], None]):
        """Generic (non-conv) gradient contribution to x: emit(dst, accumulate)."""
        if not x.requires_grad:
            return
        leaves = x.leaves()
        written = [p.grad_written for p in leaves]
        if any(written) and not all(written):
            self._zero_unwritten(x)
        acc = any(written)
        dst = self.grad_act(x)
        run = dw_free(lambda: emit(dst, acc))   # writes an activation gradient
        if hasattr(emit, "kind"):   # tagged launches (kind, algorithmic HBM bytes) keep their tag for bench.py
            _tag(run, emit.kind, emit.bytes + (4.0 * dst.n * dst.h * dst.w * dst.c if acc else 0.0), emit.label)
        self.bwd.append(run)
        for p in leaves:
            p.grad_written = True
            p.g16 = False
            p.pending -= 1

    def _bwd_conv(self, layers, xs: List[T], outs: List[T], res: List[Optional[T]]):
        G = len(layers)
        live = [i for i in range(G) if self._has_grad(outs[i])]
        for i in range(G):
            if i not in live:
                self._null(xs[i])
                if res[i] is not None:
                    self._null(res[i])
        if not live:
            return
        for i in live:
            self._finalize(outs[i])
            self._ensure_g16(outs[i])     # dY is read by the wgrad and dgrad launches below through its bf16 twin
        # weight / bias gradients
        if self.need_wgrad:
            self._emit_wgrad([(layers[i], xs[i].act, self.grad_act(outs[i])) for i in live], self.bwd, accumulate=False)
            for i in live:
                L = layers[i]
                if L.db_int is not None:
                    g = self.grad_act(outs[i])
                    if L.bias_view is not None:
                        g = Act(g.buf.view(g.n, *L.bias_view))
                    if self.defer_bias:
                        self.bias_jobs.append((L, g))
                    else:
                        self.bwd.append(lambda g=g, L=L: ops.bias_grad(g, L.db_int, True))
            for i in live:
                self.bwd_marks[layers[i].name] = len(self.bwd)
        # residual branch: its gradient is d_pre itself; defer it as an addend of the next conv contribution
        for i in live:
            r = res[i]
            if r is not None and r.requires_grad:
                r.deferred.append(self.grad_act(outs[i]))
                for p in r.leaves():
                    p.pending -= 1
        # data gradients
        dargs = []
        for i in live:
            L, x = layers[i], xs[i]
            if not x.requires_grad:
                continue
            leaves = x.leaves()
            written = [p.grad_written for p in leaves]
            if any(written) and not all(written):
                self._zero_unwritten(x)
            acc = any(written)
            dst = self.grad_act(x)
            adds: List[Act] = []
            if acc:
                adds.append(dst)
            # deferred residual addends registered on exactly this tensor ride along (max two addend slots)
            while x.deferred and len(adds) < 2:
                adds.append(x.deferred.pop())
            for p in leaves:
                p.pending -= 1
            # activation backward of the producers, fused where this is the last contribution
            slopes = []
            fuse_any = False
            for p in leaves:
                final = p.pending == 0 and not p.deferred and not x.deferred and (p.parent is None or not p.parent.deferred
                                                                               or p.parent is x and not x.deferred)
                s = LINEAR
                if final and p.slope != LINEAR and not p.masked:
                    s = p.slope
                    p.masked = True
                    fuse_any = True
                slopes.append((p, s))
            epi, slope_val, slope_vec, mask = EPI_LINEAR, 0.0, None, None
            if fuse_any:
                epi, mask = EPI_MASK, x.act
                uniq = {s for _, s in slopes}
                if len(uniq) == 1:
                    slope_val = uniq.pop()
                else:
                    vec = torch.ones(round_up(x.act.c, 4), dtype=torch.float32)
                    for p, s in slopes:
                        o = p.act.c0 - x.act.c0
                        vec[o:o + p.act.c] = s
                    slope_vec = vec.to(self.device)
                    self.keep.append(slope_vec)
            conv_written = True
            if getattr(L, "dgrad_gemm", False) and not self.exact:
                self._emit_dgrad_gemm(L, self.grad_act(outs[i]), dst, adds, mask, slope_vec, slope_val, epi)
                conv_written = False    # finished by view copies: the bf16 twin of dst is stale
            else:
                dargs.append(dict(L=L, dgrad=True, x=self.grad_act(outs[i]), out=dst,
                                  add1=adds[0] if len(adds) > 0 else None, add2=adds[1] if len(adds) > 1 else None,
                                  mask=mask, slopes=slope_vec, slope=slope_val, epilogue=epi))
            for p in leaves:
                p.grad_written = True
                p.g16 = self.bf16 and conv_written
        if dargs:
            self._emit_conv(dargs, self.bwd)

    def _emit_dgrad_gemm(self, L, dy: Act, dst: Act, adds, mask, slope_vec, slope_val, epi):
        """Input gradient of a k x k 'valid' conv whose output is 1x1 (fc1 = Linear(32768, 512) run as an 8x8 conv,
        D_and_G_model.py:212): dx[n, (tap, ci)] = sum_co dy[n, co] * Wd[tap][ci][co] is ONE GEMM (B, Cout) x (Cout, taps*Cin)
        over the existing dgrad packing read as a single [taps*Cin][Cout] matrix - every weight byte is read once, where the
        generic phase-decomposed path lets each of the 16 pixel tiles stream all 64 taps (0.35 -> ~0.03 ms).  The result goes
        through a contiguous row and is then added / masked into the (possibly strided) destination view."""
        wd = L.wd
        taps, cin = L.k * L.k, dst.c
        assert dy.h == 1 and dy.w == 1 and dst.h == L.k and dst.w == L.k and wd.rows_pad == cin and wd.rows == cin
        n = dy.n
        tmp = Act.empty(n, 1, 1, taps * cin, self.device)
        tmp8 = Act(tmp.buf.view(n, L.k, L.k, cin))
        if self.bf16:
            wd = L.wd16
        pk = ops.Packed(wd.data, 1, taps * cin, wd.k, taps * cin, wd.k_pad)
        self.keep.append((tmp, tmp8, pk))
        arg = ops.conv_args(CONV_FWD, dy, tmp, pk, 1, 1, 0, round_tf32=False, bf16=self.bf16, out16=False)
        fl = 2.0 * n * taps * cin * L.cout
        run = self._conv_launch([arg], fl, L.name + ":dgrad")
        self.bwd.append(run)
        others = [a for a in adds if a is not dst]
        acc = any(a is dst for a in adds)
        self.bwd.append(lambda: ops.view_copy(tmp8, dst, acc))
        for a in others:
            self.bwd.append(lambda a=a: ops.view_copy(a, dst, True))
        if epi == EPI_MASK:
            self.bwd.append(lambda: ops.act_backward(dst, mask, dst, slope=slope_val, slopes=slope_vec))

    # ------------------------------------------------------------------ replay
    def run_forward(self):
        for f in self.fwd:
            f()

    def run_backward(self, hooks: Optional[Dict[int, Callable[[], None]]] = None):
        if hooks:
            for i, f in enumerate(self.bwd):
                f()
                h = hooks.get(i + 1)
                if h is not None:
                    h()
        else:
            for f in self.bwd:
                f()
