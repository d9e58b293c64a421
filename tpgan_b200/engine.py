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
        # backward-trace state
        self.pending = 0
        self.grad_written = False
        self.masked = False
        self.deferred: List[Act] = []            # residual-branch addends waiting for a conv contribution

    def leaves(self) -> List["T"]:
        return self.parts if self.parts is not None else [self]

    @property
    def shape(self):
        return (self.act.n, self.act.h, self.act.w, self.act.c)


class ConvLayer:
    """One Conv2d / ConvTranspose2d of the reference model: reference-layout parameters + tensor-core packings."""

    def __init__(self, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter], transposed: bool, k: int,
                 stride: int, pad: int, name: str = ""):
        self.weight, self.bias = weight, bias
        self.transposed = transposed
        self.k, self.stride, self.pad = k, stride, pad
        self.name = name
        self.kind_fwd = DECONV_FWD if transposed else CONV_FWD
        self.kind_dgrad = DECONV_DGRAD if transposed else CONV_DGRAD
        self.cin = weight.shape[0] if transposed else weight.shape[1]
        self.cout = weight.shape[1] if transposed else weight.shape[0]
        self.ready = False

    def setup(self, in_cmap: Optional[List[int]], out_cmap: Optional[List[int]], in_c: int, out_c: int, device,
              need_dgrad: bool = True):
        """Allocate packings for the internal channel layouts seen at trace time (idempotent; layouts must not change)."""
        key = (tuple(in_cmap) if in_cmap else None, tuple(out_cmap) if out_cmap else None, in_c, out_c)
        if self.ready:
            assert key == self._key, f"{self.name}: channel layout changed between traces"
            return
        self._key = key
        mk = lambda m: None if m is None else torch.tensor(m, dtype=torch.int32, device=device)
        self.in_map, self.out_map = mk(in_cmap), mk(out_cmap)
        self.in_c, self.out_c = in_c, out_c  # internal widths (incl. padding lanes when a map is given)
        shp = tuple(self.weight.shape)
        self.wf = ops.alloc_packed(self.kind_fwd, shp, rows_int=out_c, k_int=in_c, device=device)
        self.wd = ops.alloc_packed(self.kind_dgrad, shp, rows_int=in_c, k_int=out_c, device=device) if need_dgrad else None
        self.dw = ops.alloc_packed(self.kind_fwd, shp, rows_int=out_c, k_int=in_c, device=device)
        self.bias_int = torch.zeros(round_up(out_c, 4), dtype=torch.float32, device=device) if self.bias is not None else None
        self.db_int = torch.zeros_like(self.bias_int) if self.bias is not None else None
        if out_cmap is not None:
            oc = torch.tensor(out_cmap, dtype=torch.long, device=device)
            self._b_valid = (oc >= 0).nonzero().flatten()
            self._b_ref = oc[self._b_valid]
        self.ready = True
        self.repack()

    def repack(self):
        """Reference-layout fp32 master weights -> tf32-rounded K-major packings (after every optimizer step)."""
        if not self.ready:
            return
        w = self.weight.data
        ops.pack_weights(w, self.kind_fwd, self.wf, row_map=self.out_map, k_map=self.in_map)
        if self.wd is not None:
            ops.pack_weights(w, self.kind_dgrad, self.wd, row_map=self.in_map, k_map=self.out_map)
        if self.bias is not None:
            if self.out_map is None:
                self.bias_int[: self.cout].copy_(self.bias.data)
            else:
                self.bias_int.zero_()
                self.bias_int[self._b_valid] = self.bias.data[self._b_ref]

    def zero_grad(self):
        if self.ready:
            self.dw.data.zero_()
            if self.db_int is not None:
                self.db_int.zero_()

    def export_grad(self, accumulate: bool = False):
        """Packed dW / internal db -> .grad of the reference-layout parameters."""
        if not self.ready or self.weight.grad is None:
            return
        ops.unpack_weights(self.dw, self.weight.grad, self.kind_fwd, row_map=self.out_map, k_map=self.in_map,
                           accumulate=accumulate)
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


class Plan:
    """Traced schedule for one network instance and batch size."""

    def __init__(self, device, training: bool = True, need_wgrad: bool = True):
        self.device = device
        self.training = training
        self.need_wgrad = need_wgrad
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self.tape: List[Callable[[], None]] = []
        self.grad_bufs: Dict[int, torch.Tensor] = {}
        self.keep: list = []  # keeps buffers / ctypes structs alive
        self.bytes = 0
        self.bwd_marks: Dict[str, int] = {}   # layer name -> index in self.bwd after which its dW is final

    # ------------------------------------------------------------------ buffers
    def new(self, n, h, w, c, slope=LINEAR, name="", requires_grad=True) -> T:
        a = Act.empty(n, h, w, c, self.device)
        self.bytes += a.buf.numel() * 4
        return T(a, slope, name, requires_grad)

    def wrap(self, act: Act, slope=LINEAR, name="", requires_grad=True) -> T:
        return T(act, slope, name, requires_grad)

    def concat(self, n, h, w, widths: Sequence[int], name="") -> T:
        """Concat buffer: parts are written in place by their producers (replaces torch.cat, D_and_G_model.py:100...)."""
        offs, o = [], 0
        for c in widths:
            offs.append(o)
            o += round_up(c, 4)
        buf = torch.zeros((n, h, w, o), dtype=torch.float32, device=self.device)
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
                g = torch.zeros_like(buf)
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
             residuals: Optional[Sequence[Optional[T]]] = None, name: str = "") -> List[T]:
        """Grouped conv/deconv forward (+bias, +residual, +activation).  slope None = no activation."""
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
            L.setup(x.cmap, out_cmap, x.act.c, out.act.c, self.device)
            assert (x.ref_c == L.cin) and (out.ref_c == L.cout), (L.name, x.ref_c, L.cin, out.ref_c, L.cout)
            args.append(ops.conv_args(L.kind_fwd, x.act, out.act, L.wf, L.k, L.stride, L.pad, bias=L.bias_int,
                                      add1=None if r is None else r.act, slope=0.0 if slope is None else slope,
                                      epilogue=EPI_LINEAR if slope is None else EPI_LEAKY))
            self.use(x)
            if r is not None:
                self.use(r)
            outs_l.append(out)
        self.fwd.append(self._conv_launch(args))
        self.tape.append(lambda: self._bwd_conv(layers, list(xs), outs_l, res))
        return outs_l

    def _conv_launch(self, args):
        arr = (ops.ConvArgs * len(args))(*args)
        self.keep.append(arr)
        n = len(args)
        lib = ops._lib.load()

        def run():
            ops._lib.check(lib.tpgan_conv2d(arr, n, ops._stream()), "conv2d")
        return run

    def _wgrad_launch(self, args):
        arr = (ops.WgradArgs * len(args))(*args)
        self.keep.append(arr)
        n = len(args)
        lib = ops._lib.load()

        def run():
            ops._lib.check(lib.tpgan_conv2d_wgrad(arr, n, ops._stream()), "wgrad")
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

    # ------------------------------------------------------------------ backward tracing
    def seed_grad(self, t: T):
        """Declare that an external kernel (a loss) writes t's gradient before the backward replay."""
        self.grad_act(t)
        for p in t.leaves():
            p.grad_written = True

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

    def _zero_unwritten(self, t: T):
        for p in t.leaves():
            if not p.grad_written:
                g = self.grad_act(p)
                self.bwd.append(lambda g=g: g.buf[..., g.c0:g.c0 + g.c].zero_())
                p.grad_written = True

    def _finalize(self, t: T):
        """Make t.grad hold d(loss)/d(pre-activation): flush deferred addends, apply the activation mask if nobody fused
        it."""
        self._flush_deferred(t)
        for p in t.leaves():
            if p.slope != LINEAR and not p.masked and p.grad_written:
                g = self.grad_act(p)
                self.bwd.append(lambda g=g, p=p: ops.act_backward(g, p.act, g, slope=p.slope))
                p.masked = True

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
        self.bwd.append(lambda: emit(dst, acc))
        for p in leaves:
            p.grad_written = True
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
        # weight / bias gradients
        if self.need_wgrad:
            wargs = [ops.wgrad_args(layers[i].kind_fwd, xs[i].act, self.grad_act(outs[i]), layers[i].dw, layers[i].k,
                                    layers[i].stride, layers[i].pad) for i in live]
            self.bwd.append(self._wgrad_launch(wargs))
            for i in live:
                L = layers[i]
                if L.db_int is not None:
                    g = self.grad_act(outs[i])
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
            dargs.append(ops.conv_args(L.kind_dgrad, self.grad_act(outs[i]), dst, L.wd, L.k, L.stride, L.pad,
                                       add1=adds[0] if len(adds) > 0 else None, add2=adds[1] if len(adds) > 1 else None,
                                       mask=mask, slopes=slope_vec, slope=slope_val, epilogue=epi, round_tf32=True))
            for p in leaves:
                p.grad_written = True
        if dargs:
            self.bwd.append(self._conv_launch(dargs))

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
