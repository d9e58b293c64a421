"""The TP-GAN G+D training step on the C-ABI kernels.

The reference ships no training loop and no loss code - only the loss weights (config.py:71-82), `train` hyper-parameters
(config.py:50-57), set_requires_grad (UtilityMethods.py:43-56) and the TrainDataset batch keys (DataAndDataset.py:206-226).
The step implemented here is the one oracle/step.py defines from those and is checked against it:

  D phase : L_D = mean D(fake) - mean D(real) + 10 * mean_n (||grad_xhat sum D(xhat_n)||_2 - 1)^2     (WGAN-GP critic)
  G phase : L_G = 1.0*pixel + 3.0*local + 0.3*symmetry + 1e-3*(-mean D(fake)) + 1e-3*tv + 10*CE
  Adam(lr = train['learning_rate'] = 1e-4) on both networks.

Everything on the device is a launch of libtpgan_b200.so:
  * landmark crop of the 4 profile + 4 frontal patches (one kernel each),
  * the generator forward/backward plan of tpgan_b200.D_and_G_model (tcgen05 convs, grouped local pathways, fused stitch),
  * the critic: ONE batched forward over [fake; real; xhat], one batched dgrad chain, weight gradients on the fake/real
    part, and the gradient penalty's double backward as a *tangent forward* - because D is piecewise linear
    (conv + LeakyReLU, no BN), d GP / d W_l = wgrad(x = v_{l-1}, dy = e_l) where e_l are the masked deltas of the
    backward pass seeded with ones and v_l = mask_l * (W_l v_{l-1}) is the forward pass of u = dGP/dg through the same
    masks.  No autograd graph, no second-order kernels.
  * fused image losses (pixel L1 at 3 scales + symmetry + TV in one pass, writing dL/dfake), patch L1, softmax-CE,
  * flat-buffer Adam (one launch per network), weight re-packing into the tensor-core layout.
Data parallelism (tpgan_b200.parallel) all-reduces the flat gradient buffers in buckets overlapped with backward.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import config as cfg
from . import ops
from .D_and_G_model import PART_NAMES, PATCH_HW, Discriminator, Generator, _layer, _unpack_conv_seq
from .engine import LINEAR, ConvLayer, GradArena, Plan, T, dw_free
from .ModificationLayer import ResidualBlock, _negative_slope
from .ops import Act, EPI_LEAKY, EPI_LINEAR, EPI_MASK


def _sl(a: Act, n0: int, n1: int) -> Act:
    """Batch slice [n0, n1) of an activation (same channel window)."""
    return Act(a.buf[n0:n1], a.c0, a.c)


# ---------------------------------------------------------------------------------------------------------- flat buffers
class FlatParams:
    """All parameters of a module as views into ONE flat fp32 buffer (and their .grad into a second one), so the optimizer
    is a single kernel launch and the gradient all-reduce works on contiguous buckets.  `order` = parameter names in the
    order their gradients become final during backward (bucket order)."""

    def __init__(self, module: torch.nn.Module, order: Optional[Sequence[str]] = None):
        named = dict(module.named_parameters())
        names = list(named.keys())
        if order is not None:
            seen = set()
            names = [n for n in order if n in named and not (n in seen or seen.add(n))] + \
                    [n for n in named if n not in set(order)]
        self.names = names
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in names:
            self.offsets[n] = off
            off += ops.round_up(named[n].numel(), 4)
        self.total = off
        dev = named[names[0]].device
        self.data = ops.zeros(off, torch.float32, dev)
        self.grad = ops.zeros(off, torch.float32, dev)
        self.m = ops.zeros(off, torch.float32, dev)
        self.v = ops.zeros(off, torch.float32, dev)
        self.step = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        for n in names:
            p = named[n]
            o, k = self.offsets[n], p.numel()
            self.data[o:o + k].copy_(p.data.flatten())
            p.data = self.data[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)

    def adam(self, lr: float, grad_scale: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        """torch.optim.Adam update of the whole flat buffer; the step count is device-resident (graph-capturable)."""
        self.step += 1
        ops.adam_step_dev(self.data, self.grad, self.m, self.v, lr, betas[0], betas[1], eps, weight_decay, self.step_dev,
                          grad_scale)


class FlatOptimizer:
    """torch.optim-compatible view of the fused optimizer state kept in a FlatParams (momentum / Adam moments as flat
    buffers, step count on the device), so that the reference's checkpoint helpers work unchanged:
    `save_optimizer(trainer.optimizer, model, dir, epoch)` (UtilityMethods.py:78-103) writes the same
    {'optimizer', 'model', 'epoch'} file a torch.optim.SGD / Adam over `module.parameters()` would, and a file written
    by the reference's loop loads back with load_state_dict()."""

    def __init__(self, flat: "FlatParams", module: torch.nn.Module, kind: str, hyper: Dict[str, float],
                 lr_get: Optional[Callable[[], float]] = None, lr_set: Optional[Callable[[float], None]] = None):
        """lr_get / lr_set: accessors of the LIVE learning rate a trainer's scheduler decays (PretrainTrainer.end_epoch keeps
        it on the host and in device memory); when given, state_dict() records it (and initial_lr, as torch's MultiStepLR
        does) and load_state_dict() restores it."""
        assert kind in ("sgd", "adam")
        self.flat, self.module, self.kind, self.hyper = flat, module, kind, dict(hyper)
        self.lr_get, self.lr_set = lr_get, lr_set

    def _template_group(self) -> dict:
        p = [torch.nn.Parameter(torch.zeros(1))]
        opt = torch.optim.SGD(p, **self.hyper) if self.kind == "sgd" else torch.optim.Adam(p, **self.hyper)
        return dict(opt.state_dict()["param_groups"][0])

    def _slices(self):
        for i, (name, p) in enumerate(self.module.named_parameters()):
            o = self.flat.offsets[name]
            yield i, p, slice(o, o + p.numel())

    def state_dict(self) -> dict:
        state = {}
        steps = int(self.flat.step_dev.item()) if self.kind == "adam" else 0
        for i, p, sl in self._slices():
            if self.kind == "sgd":
                state[i] = {"momentum_buffer": self.flat.m[sl].view(p.shape).clone()}
            else:
                state[i] = {"step": torch.tensor(float(steps)), "exp_avg": self.flat.m[sl].view(p.shape).clone(),
                            "exp_avg_sq": self.flat.v[sl].view(p.shape).clone()}
        group = self._template_group()
        group["params"] = list(range(len(state)))
        if self.lr_get is not None:
            group["initial_lr"] = group["lr"]
            group["lr"] = float(self.lr_get())
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict):
        n = sum(1 for _ in self.module.parameters())
        ids = sd["param_groups"][0]["params"]
        if len(sd["param_groups"]) != 1 or len(ids) != n:
            raise ValueError("optimizer state does not match the module's parameter list")
        if self.lr_set is not None and "lr" in sd["param_groups"][0]:
            self.lr_set(float(sd["param_groups"][0]["lr"]))
        for i, p, sl in self._slices():
            st = sd["state"].get(ids[i], sd["state"].get(str(ids[i])))
            if st is None:                       # torch omits parameters that never received a gradient
                self.flat.m[sl].zero_()
                self.flat.v[sl].zero_()
                continue
            if self.kind == "sgd":
                self.flat.m[sl].copy_(st["momentum_buffer"].flatten())
            else:
                self.flat.m[sl].copy_(st["exp_avg"].flatten())
                self.flat.v[sl].copy_(st["exp_avg_sq"].flatten())
                self.flat.step_dev.fill_(int(st["step"]))
                self.flat.step = int(st["step"])


class LayerSet:
    """Re-packing (after the optimizer) and gradient export (after backward) of a set of layers as multi-tensor launches:
    one pack, one transpose and one unpack launch for all eligible layers, individual calls for the few special ones
    (fc1's 128 KB rows, the deconv-as-GEMM layer)."""

    def __init__(self, layers: Sequence[ConvLayer], device, bias_jobs: Sequence = ()):
        self.multi = [L for L in layers if L.multi_ok()]
        self.single = [L for L in layers if not L.multi_ok()]
        packs, ptr, utr, unp = [], [], [], []
        for L in self.multi:
            a, b = L.pack_jobs()
            packs += a
            ptr += b
            c, d = L.unpack_jobs(False)
            utr += c
            unp += d
        casts = [j for L in self.multi for j in L.cast_jobs()]
        self.cast_tab = ops.JobTable("cast", casts, device)      # bf16 operand copies of the packings (bf16 mode only)
        self.pack_tab = ops.JobTable("pack", packs, device)
        self.ptrans_tab = ops.JobTable("transpose", ptr, device)
        self.utrans_tab = ops.JobTable("transpose", utr, device)
        self.unpack_tab = ops.JobTable("pack", unp, device, unpack=True)
        ids = {id(L) for L in layers}
        self.bias_tab = ops.JobTable("bias", [ops.bias_job(g, L.db_int) for L, g in bias_jobs if id(L) in ids], device)

    def repack(self):
        self.pack_tab.run()
        self.ptrans_tab.run()
        self.cast_tab.run()
        for L in self.single:
            L.repack()

    def export(self):
        self.bias_tab.run()
        self.utrans_tab.run()
        self.unpack_tab.run()
        for L in self.single:
            L.export_grad(accumulate=False)


# ---------------------------------------------------------------------------------------------------------- CUDA graphs
class Eager:
    """Marks a schedule entry that must run eagerly (host-visible side effects)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        self.fn()


class Collective(Eager):
    """A stream-ordered NCCL collective (possibly forked onto a communication stream with events).  Runs eagerly between
    graph segments by default; with GraphRunner(capture_collectives=True) it is captured INTO the graph like any other
    launch (torch's NCCL process group is capturable), so a data-parallel step is a single graph replay."""


class GraphRunner:
    """Replays a schedule (list of callables) as CUDA graphs: every maximal run of capturable entries becomes one graph,
    `Eager` entries run between them.  All device pointers in a schedule are static (pre-allocated plans)."""

    def __init__(self, schedule: Sequence[Callable], capture_collectives: bool = False):
        self.segments: list = []
        run: List[Callable] = []
        for f in schedule:
            if isinstance(f, Eager) and not (capture_collectives and isinstance(f, Collective)):
                if run:
                    self.segments.append(run)
                    run = []
                self.segments.append(f)
            else:
                run.append(f)
        if run:
            self.segments.append(run)
        self.graphs: Optional[list] = None
        self.warm = False
        self.kernels_per_run = 0
        import os
        # weight gradients on a parallel graph branch (see _capture_two_streams); TPGAN_SIDE_WGRAD=0 captures one stream
        self.side_wgrad = os.environ.get("TPGAN_SIDE_WGRAD", "1") != "0"
        self._side = torch.cuda.Stream() if self.side_wgrad else None
        # data parallel: SMs left to NCCL by the launches right behind an overlapped all-reduce (see capture())
        self.reserve_sms = int(os.environ.get("TPGAN_DP_RESERVE", "0"))
        self.reserve_launches = int(os.environ.get("TPGAN_DP_RESERVE_LAUNCHES", "3"))

    def run_eager(self):
        for seg in self.segments:
            if isinstance(seg, Eager):
                seg()
            else:
                for f in seg:
                    f()

    def capture(self):
        from . import _lib
        torch.cuda.synchronize()
        self.graphs = []
        l0 = _lib.launch_count()
        after_collective = False
        for seg in self.segments:
            if isinstance(seg, Eager):
                seg()
                self.graphs.append(seg)
                after_collective = isinstance(seg, Collective)
                continue
            # A segment that starts right behind an overlapped all-reduce: its first launches are captured with a few SMs left
            # free, so that NCCL's CTAs run NEXT to the persistent kernels instead of delaying the CTAs that would have
            # taken those SMs (a statically partitioned launch is as slow as its last CTA).  Only those launches pay.
            window = self.reserve_launches if (after_collective and self.reserve_sms > 0) else 0
            if window:
                seg = self._with_reserve(seg, window)
            after_collective = False
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if self.side_wgrad:
                    self._capture_two_streams(seg)
                else:
                    for f in seg:
                        f()
            g.replay()          # capture does not execute: run the segment once so later segments see its results
            self.graphs.append(g)
        self.kernels_per_run = _lib.launch_count() - l0   # kernels of this library recorded into the graphs
        torch.cuda.synchronize()

    def _with_reserve(self, seg, window: int):
        """seg with tpgan_set_sm_reserve(reserve_sms) around its first `window` tensor-core launches (host-side planning
        state: it shapes the grid of the launches recorded while it is set, so it must be set at CAPTURE time)."""
        from . import _lib
        out, seen, on = [], 0, False
        for f in seg:
            is_tc = getattr(f, "kind", None) in ("tapgemm", "rowconv", "wgrad")
            if is_tc and seen < window:
                def wrapped(f=f):
                    prev = _lib.set_sm_reserve(self.reserve_sms)
                    try:
                        f()
                    finally:
                        _lib.set_sm_reserve(prev)
                for attr in ("kind", "flops", "label", "dw_free"):
                    if hasattr(f, attr):
                        setattr(wrapped, attr, getattr(f, attr))
                out.append(wrapped)
                seen += 1
            else:
                out.append(f)
        return out

    def _capture_two_streams(self, seg):
        """Weight-gradient launches go to a side stream (a parallel branch of the captured graph): wgrad of layer l needs only
        dy_l - final before the launch is reached - and the stored forward activation, and nothing later in the backward pass
        writes either, so it may overlap the input-gradient chain that continues on the main stream; its dW is read by the
        export / optimizer launches only.  The main stream joins the branch before every launch that is not itself a
        convolution (anything that might read dW) and at the end of the segment."""
        main = torch.cuda.current_stream()
        side = self._side
        pending = False
        for f in seg:
            kind = getattr(f, "kind", None)
            if kind in ("wgrad", "dw_wgrad"):      # dense and depthwise weight gradients
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    f()
                pending = True
                continue
            if pending and kind not in ("tapgemm", "rowconv", "cast16", "dw_dgrad", "dw_fwd") and not getattr(f, "dw_free", False):
                ev = torch.cuda.Event()
                ev.record(side)
                main.wait_event(ev)
                pending = False
            f()
        if pending:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

    def run(self):
        if not self.warm:       # first call: plain launches (CUDA lazily loads kernels on first use, which capture forbids)
            self.run_eager()
            self.warm = True
            return
        if self.graphs is None:
            self.capture()      # the capturing call also executes the schedule once
            return
        for g in self.graphs:
            if isinstance(g, Eager):
                g()
            else:
                g.replay()


# ---------------------------------------------------------------------------------------------------------- critic
class CriticPlan:
    """Static schedule of the discriminator (reference D_and_G_model.py:409-435) for a batch of M images, with the native
    WGAN-GP machinery described in the module docstring."""

    def __init__(self, D: Discriminator, M: int, B: int, device, exact: bool = False, defer_pack: bool = False,
                 bf16: bool = False):
        self.D, self.M, self.B, self.device, self.exact = D, M, B, device, exact
        self.defer_pack = defer_pack
        self.bf16 = bf16
        self.h = Plan(device, exact=exact, bf16=bf16)  # launch emitters / keep-alive only
        new = lambda n, hw, c: Act.empty(n, hw, hw, c, device)
        self.x0 = new(M, 128, 3)
        self.ops_: List[dict] = []
        cur, hw = self.x0, 128
        self.layers: List[ConvLayer] = []
        for i, m in enumerate(D.model):
            if isinstance(m, ResidualBlock):
                assert len(m.shortcut) == 0 and m.scaling_factor == 1.0
                L1 = self._mk(m.layers[0], f"model.{i}.layers.0", cur.c)
                L2 = self._mk(m.layers[1], f"model.{i}.layers.1", cur.c)
                hmid, y = new(M, hw, L1.cout), new(M, hw, L2.cout)
                s1 = _unpack_conv_seq(m.layers[0])[2]
                self.ops_.append(dict(kind="res", L1=L1, L2=L2, x=cur, h=hmid, y=y, s1=s1, s2=_negative_slope(m.activation)))
                cur = y
            else:
                _, cm, slope, _bn = _unpack_conv_seq(m)
                L = self._mk(m, f"model.{i}", cur.c)
                hw = (hw + 2 * L.pad - L.k) // L.stride + 1
                y = new(M, hw, L.cout)
                self.ops_.append(dict(kind="conv", L=L, x=cur, y=y, s=slope))
                cur = y
        self.logits = cur
        # gradient (masked delta) buffers, one per activation, and tangent buffers (B images)
        self.g: Dict[int, Act] = {}
        self.v: Dict[int, Act] = {}
        for op in self.ops_:
            for a in ([op["x"], op["y"]] + ([op["h"]] if op["kind"] == "res" else [])):
                if id(a.buf) not in self.g:
                    self.g[id(a.buf)] = Act.empty(M, a.h, a.w, a.c, device)
                    self.v[id(a.buf)] = Act.empty(B, a.h, a.w, a.c, device)
        self.g_logits_g = Act.empty(B, self.logits.h, self.logits.w, 1, device)  # seed of the G phase
        self.sq = torch.zeros(B, dtype=torch.float32, device=device)
        self.coeff = torch.zeros(B, dtype=torch.float32, device=device)
        self.gp_sum = torch.zeros(1, dtype=torch.float32, device=device)
        self._slope_of: Dict[int, float] = {id(self.x0.buf): LINEAR}
        for op in self.ops_:
            if op["kind"] == "res":
                self._slope_of[id(op["h"].buf)] = op["s1"]
                self._slope_of[id(op["y"].buf)] = op["s2"]
            else:
                self._slope_of[id(op["y"].buf)] = LINEAR if op["s"] is None else op["s"]
        self._build()

    def _mk(self, seq, name, in_c) -> ConvLayer:
        _, cm, _, _bn = _unpack_conv_seq(seq)
        L = _layer(cm, name)
        L.setup(None, None, in_c, L.cout, self.device, exact=self.exact, defer_pack=self.defer_pack, bf16=self.bf16)
        self.layers.append(L)
        return L

    def G_(self, a: Act) -> Act:
        return self.g[id(a.buf)]

    def V_(self, a: Act) -> Act:
        return self.v[id(a.buf)]

    # ---- schedule construction
    def _fwd_list(self, n0, n1) -> List[Callable]:
        lst: List[Callable] = []
        e = self.h._emit_conv
        for op in self.ops_:
            if op["kind"] == "conv":
                L, s = op["L"], op["s"]
                e([dict(L=L, dgrad=False, x=_sl(op["x"], n0, n1), out=_sl(op["y"], n0, n1), bias=L.bias_int,
                        slope=0.0 if s is None else s, epilogue=EPI_LINEAR if s is None else EPI_LEAKY)], lst)
            else:
                L1, L2 = op["L1"], op["L2"]
                e([dict(L=L1, dgrad=False, x=_sl(op["x"], n0, n1), out=_sl(op["h"], n0, n1), bias=L1.bias_int,
                        slope=op["s1"], epilogue=EPI_LEAKY)], lst)
                e([dict(L=L2, dgrad=False, x=_sl(op["h"], n0, n1), out=_sl(op["y"], n0, n1), bias=L2.bias_int,
                        add1=_sl(op["x"], n0, n1), slope=op["s2"], epilogue=EPI_LEAKY)], lst)
        return lst

    def _dgrad(self, lst, L, dy: Act, x_full: Act, n0: int, n1: int, add: Optional[Act]):
        """G(x)[n0:n1] = mask_x * (W^T dy + add)."""
        sx = self._slope_of[id(x_full.buf)]
        kw = dict(L=L, dgrad=True, x=dy, out=_sl(self.G_(x_full), n0, n1), add1=add)
        if sx != LINEAR:
            kw.update(mask=_sl(x_full, n0, n1), slope=sx, epilogue=EPI_MASK)
        self.h._emit_conv([kw], lst)

    def _bwd_list(self, n0, n1, w0, w1, g_logits: Act, dx_out: Optional[Act], dx_acc: bool, x0r=None) -> List[Callable]:
        """dgrad chain over images [n0,n1) seeded from g_logits; weight/bias gradients over [w0,w1) (w1<=w0: none);
        the input gradient of images x0r=[a,b) goes to dx_out (accumulating when dx_acc) or is skipped when None."""
        x0r = (n0, n1) if x0r is None else x0r
        lst: List[Callable] = []
        sl = lambda a: _sl(a, n0, n1)
        gout = lambda a: g_logits if a is self.logits else sl(self.G_(a))

        def wg(L, x, dy_full):
            if w1 > w0:
                self.h._emit_wgrad([(L, _sl(x, w0, w1), dy_full)], lst)
                if L.db_int is not None:
                    lst.append(dw_free(lambda d=dy_full, L=L: ops.bias_grad(d, L.db_int, True)))

        def dyw(a):  # delta of activation `a` restricted to the weight-gradient images
            if a is self.logits:
                return _sl(g_logits, w0 - n0, w1 - n0)
            return _sl(self.G_(a), w0, w1)

        for op in reversed(self.ops_):
            if op["kind"] == "conv":
                L, x, y = op["L"], op["x"], op["y"]
                wg(L, x, dyw(y) if w1 > w0 else None)
                if x is self.x0:
                    if dx_out is not None:
                        dy0 = _sl(self.G_(y), x0r[0], x0r[1])
                        self.h._emit_conv([dict(L=L, dgrad=True, x=dy0, out=dx_out, add1=dx_out if dx_acc else None)], lst)
                else:
                    self._dgrad(lst, L, gout(y), x, n0, n1, None)
            else:
                L1, L2, x, hm, y = op["L1"], op["L2"], op["x"], op["h"], op["y"]
                wg(L2, hm, dyw(y) if w1 > w0 else None)
                self._dgrad(lst, L2, gout(y), hm, n0, n1, None)
                wg(L1, x, dyw(hm) if w1 > w0 else None)
                self._dgrad(lst, L1, sl(self.G_(hm)), x, n0, n1, gout(y))
        return lst

    def _tangent_list(self, n0, n1) -> List[Callable]:
        """v_l = mask_l * (W_l v_{l-1} [+ v_skip]) over the masks of images [n0,n1), then dW_l += wgrad(v_{l-1}, e_l)."""
        lst: List[Callable] = []
        e = self.h._emit_conv
        sl = lambda a: _sl(a, n0, n1)

        def tconv(L, vin, a_out, add=None, last=False):
            self.h._emit_wgrad([(L, vin, sl(self.G_(a_out)))], lst)
            if last:
                return
            s = self._slope_of[id(a_out.buf)]
            kw = dict(L=L, dgrad=False, x=vin, out=self.V_(a_out), add1=add)
            if s != LINEAR:
                kw.update(mask=sl(a_out), slope=s, epilogue=EPI_MASK)
            e([kw], lst)

        n_ops = len(self.ops_)
        for i, op in enumerate(self.ops_):
            if op["kind"] == "conv":
                tconv(op["L"], self.V_(op["x"]), op["y"], last=(i == n_ops - 1))
            else:
                tconv(op["L1"], self.V_(op["x"]), op["h"])
                tconv(op["L2"], self.V_(op["h"]), op["y"], add=self.V_(op["x"]))
        return lst

    def _build(self):
        B, M = self.B, self.M
        self.fwd_all = self._fwd_list(0, M)
        self.fwd_fake = self._fwd_list(0, B)
        if M >= 3 * B:
            gl = self.G_(self.logits)
            # D phase: deltas for all 3B images; weight grads from fake+real; input gradient only for xhat
            self.g_x = Act.empty(B, 128, 128, 3, self.device)   # d sum D(xhat) / d xhat
            self.bwd_d = self._bwd_list(0, M, 0, 2 * B, gl, self.g_x, False, x0r=(2 * B, 3 * B))
            self.tangent = self._tangent_list(2 * B, 3 * B)
        self.dx_g: Optional[Act] = None

    def build_g_phase(self, dfake: Act):
        """G phase: dgrad-only chain over the fake images, input gradient accumulated into `dfake`."""
        self.dx_g = dfake
        self.bwd_g = self._bwd_list(0, self.B, 0, 0, self.g_logits_g, dfake, True)

    # ---- execution
    @staticmethod
    def _run(lst):
        for f in lst:
            f()

    def zero_grad(self):
        GradArena.get(self.device).zero()   # clears every layer's packed gradient accumulators (G's are stale here)

    def repack(self):
        for L in self.layers:
            L.repack()

    def export_grads(self):
        for L in self.layers:
            L.export_grad(accumulate=False)

    def d_phase_list(self, gp_weight: float) -> List[Callable]:
        """x0 must hold [fake; real; xhat].  Leaves dW/db of L_D in the layers' packed gradient buffers."""
        B = self.B
        gl = self.G_(self.logits)
        c = 1.0 / (B * self.logits.h * self.logits.w)
        gx = self.g_x
        pre = [lambda: ops.fill(_sl(gl, 0, B), c), lambda: ops.fill(_sl(gl, B, 2 * B), -c),
               lambda: ops.fill(_sl(gl, 2 * B, 3 * B), 1.0), self.zero_grad]
        mid = [dw_free(lambda: self.gp_sum.zero_()), dw_free(lambda: ops.sample_sqnorm(gx, self.sq)),
               dw_free(lambda: ops.gp_coeff(self.sq, self.coeff, gp_weight * 2.0 / B, self.gp_sum)),
               dw_free(lambda: ops.sample_scale(gx, self.coeff, self.V_(self.x0)))]
        if self.bf16:   # bf16 twins of the tensors the convs read that pointwise kernels wrote: x0 (the caller filled it),
            # the seeded logit deltas, the tangent seed; every other operand is written by a conv epilogue together with its twin
            x0, v0 = self.x0, self.V_(self.x0)
            pre += [lambda: ops.cast_bf16(x0), lambda: ops.cast_bf16(gl)]
            mid += [dw_free(lambda: ops.cast_bf16(v0))]
        return pre + self.fwd_all + self.bwd_d + mid + self.tangent

    def g_phase_list(self, adv_weight: float) -> List[Callable]:
        """D(fake) with the current weights, then d(-adv_weight * mean D(fake))/d fake accumulated into dfake."""
        B = self.B
        seed = -adv_weight / (B * self.logits.h * self.logits.w)
        pre = [lambda: ops.fill(self.g_logits_g, seed)]
        if self.bf16:   # x0[:B] still holds the fake images and their twins from the D phase
            pre.append(lambda: ops.cast_bf16(self.g_logits_g))
        return pre + self.fwd_fake + self.bwd_g

    def d_phase(self, gp_weight: float):
        self._run(self.d_phase_list(gp_weight))

    def g_phase(self, adv_weight: float):
        self._run(self.g_phase_list(adv_weight))


# ---------------------------------------------------------------------------------------------------------- identity
class IdentityPlan:
    """Identity-preserving loss (config.py:79 weight_identity_preserving; TP-GAN eq. 5) through a FROZEN feature network
    (FeatureExtract.py:5-41 / ResNet.py:5-119, restated - see tpgan_b200/ResNet.py): features of the frontal ground
    truth (no gradient) and of the fake image, L1 on the two last feature layers (pooled 512 and FC0), input gradient
    of the fake branch accumulated into d fake.  No weight gradients: the network is frozen."""

    def __init__(self, net, B: int, fake: Act, dfake: Act, frontal: Act, weight: float, sums: torch.Tensor, device,
                 exact: bool = False, bf16: bool = False):
        from .FeatureExtract import FeatureExtractModel
        base = net.base_model if isinstance(net, FeatureExtractModel) else net
        assert not base.training, "the identity network must be in eval() mode (frozen, BatchNorm folded)"
        self.base = base
        # fake branch: forward + input gradient
        pf = Plan(device, training=True, need_wgrad=False, exact=exact, bf16=bf16)
        xin = pf.wrap(fake, name="fake", requires_grad=True, s16=bf16)   # fake's twin: written by G's last conv epilogue
        xin.grad = dfake
        xin.grad_written = True           # image / adversarial terms are already in d fake: accumulate
        pooled, fc0, _ = base.trace(pf, xin, with_logits=False)
        # gt branch: forward only
        pg = Plan(device, training=False, need_wgrad=False, exact=exact, bf16=bf16)
        gin = pg.wrap(frontal, name="frontal", requires_grad=False)
        pooled_gt, fc0_gt, _ = base.trace(pg, gin, with_logits=False)
        feats = [(pooled, pooled_gt)] + ([(fc0, fc0_gt)] if fc0 is not None else [])
        self.feats = [a for a, _ in feats]
        self.loss: List[Callable] = []
        self.denoms: List[float] = []
        for i, (a, b) in enumerate(feats):
            pf.seed_grad(a)
            n = float(B * a.act.c)
            self.denoms.append(n)
            self.loss.append(lambda a=a, b=b, i=i, n=n: ops.l1_loss(a.act, b.act, pf.grad_act(a), weight / n, sums[i:i + 1]))
        pf.trace_backward()
        self.pf, self.pg = pf, pg
        self.sums = sums

    def schedule(self) -> List[Callable]:
        return self.pg.fwd + self.pf.fwd + self.loss + self.pf.bwd

    def value(self, s: Sequence[float]) -> float:
        return sum(v / n for v, n in zip(s, self.denoms))


# ---------------------------------------------------------------------------------------------------------- trainer
class TPGANTrainer:
    """One object = both networks, their plans for per-GPU batch `B`, flat parameter buffers and Adam state.

    step(batch) consumes NCHW CUDA tensors with the TrainDataset conventions (DataAndDataset.py:206-226):
      img, img_frontal (B,3,128,128) in [-1,1]; img64_frontal, img32_frontal; landmarks (B,5,2) float32 (x,y);
      z (B,64); label (B,) int64; gp_alpha (B,) float32 (the WGAN-GP interpolation coefficients).
    Patches are cropped on the device from img / img_frontal with the reference's process() arithmetic
    (DataAndDataset.py:10-56)."""

    def __init__(self, G: Generator, D: Discriminator, B: int, device="cuda", use_dropout: bool = False,
                 exact: bool = False, world_size: int = 1, group=None, bucket_mb: float = 128.0, use_graphs: bool = False,
                 identity_net=None, input_format: str = "float", dtype: str = "tf32", overlap_allreduce: bool = True,
                 graph_collectives: bool = False, force_reducer: bool = False):
        """identity_net: optional frozen FeatureExtractModel / ResNet18 in eval() mode; adds the identity-preserving
        term weight_identity_preserving * L_ip to the generator loss.
        dtype: "tf32" (BASELINE configs[1]) or "bf16" (configs[2]): bf16 tensor-core operands (activations, activation
        gradients and weights rounded to bf16 where a convolution reads them), fp32 accumulation, fp32 master weights,
        fp32 residual / gradient accumulation and losses."""
        assert dtype in ("tf32", "bf16")
        self.dtype, self.bf16 = dtype, dtype == "bf16"
        # data parallelism: overlap_allreduce = bucketed all-reduce of G's gradients on a side stream during backward (else
        # one all-reduce after backward); graph_collectives = capture the NCCL calls into the step's CUDA graph (one replay
        # per step instead of one graph segment per bucket); force_reducer = build the bucket reducer on a 1-rank group (tests)
        self.overlap_allreduce, self.graph_collectives, self.force_reducer = overlap_allreduce, graph_collectives, force_reducer
        assert not (self.bf16 and exact), "exact is the tf32 verification mode"
        self.G, self.D, self.B, self.device = G, D, B, torch.device(device)
        has_bn = lambda net: any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in net.modules())
        if has_bn(D):
            raise NotImplementedError("the fused step is built for a BatchNorm-free critic (config.py:68; the gradient penalty's "
                                      "double backward through batch statistics is not built); a Discriminator(use_batchnorm="
                                      "True) trains through the module API (forward / backward / optimizer)")
        # Generator(use_batchnorm=True) - the reference constructor's default (D_and_G_model.py:351; config.py:63 turns it
        # off): conv / deconv (no bias) -> batch-statistics BatchNorm2d -> (Leaky)ReLU in every factory stack.  The BatchNorm
        # kernels write their affine gradients straight into the flat gradient buffer and update the running statistics in
        # place; one generator forward per step = one statistics update per step, as in the oracle step.
        self.g_bn = has_bn(G)
        if self.g_bn:
            G.train()
            self.overlap_allreduce = self.force_reducer = False   # one all-reduce after backward (no bucket order for BatchNorm)
        self.steps = 0
        self.use_dropout, self.exact = use_dropout, exact
        # "float": TrainDataset tensors (img, img_frontal, img64_frontal, img32_frontal as NCHW fp32 in [-1,1]);
        # "uint8": raw HWC bytes img_u8 / img_frontal_u8 (B,128,128,3) - ToTensor()*2-1 and the 64/32 targets are computed
        # on the device (tpgan_u8_to_nhwc, tpgan_pyramid): 4.6x fewer host->device bytes per step
        assert input_format in ("float", "uint8")
        self.input_format = input_format
        self.INPUT_KEYS = self.INPUT_KEYS_U8 if input_format == "uint8" else self.INPUT_KEYS
        self.w = dict(cfg.loss)
        self.lr = cfg.train["learning_rate"]
        self.world_size, self.group = world_size, group
        self.use_graphs = use_graphs
        # every buffer of the trainer (activations, gradients, packed weights, flat parameter / optimizer state) is carved out
        # of a few large zeroed chunks: one fill launch per 256 MB instead of one per buffer
        self.arena = ops.Arena(self.device, shadow=self.bf16)
        with ops.use_arena(self.arena):
            self._init(G, D, B, exact, world_size, group, bucket_mb, identity_net)

    def _init(self, G, D, B, exact, world_size, group, bucket_mb, identity_net):
        if self.g_bn:
            self.flat_g = FlatParams(G)    # before tracing: the BatchNorm backward launches capture their .grad targets
        self._build_g()
        self.critic = CriticPlan(D, 3 * B, B, self.device, exact=exact, defer_pack=not exact, bf16=self.bf16)
        self.critic.build_g_phase(self.plan.grad_act(self.fake))
        if not self.g_bn:
            self.flat_g = FlatParams(G, self._ready_order())
        self.flat_d = FlatParams(D)
        adam = dict(lr=self.lr)      # torch.optim.Adam defaults otherwise, as FlatParams.adam
        self.optimizer_g, self.optimizer_d = FlatOptimizer(self.flat_g, G, "adam", adam), FlatOptimizer(self.flat_d, D, "adam", adam)
        self.g_set = LayerSet(self.plan.layers, self.device, self.plan.bias_jobs)
        self.d_set = LayerSet(self.critic.layers, self.device)
        self.g_set.repack()
        self.d_set.repack()
        self.reducer = None
        if (world_size > 1 and self.overlap_allreduce) or self.force_reducer:
            from .parallel import BucketReducer
            self.reducer = BucketReducer(self, bucket_mb, group)
        self.sums = torch.zeros(16, dtype=torch.float32, device=self.device)  # 0..7 image terms, 8..11 local parts, 12 ce, 13..14 identity
        self.identity = None
        if identity_net is not None:
            self.identity = IdentityPlan(identity_net, B, self.fake.act, self.plan.grad_act(self.fake), self.frontal,
                                         float(self.w["weight_identity_preserving"]), self.sums[13:15], self.device, exact=exact,
                                         bf16=self.bf16)
        self._d_logits = torch.zeros_like(self.critic.logits.buf[:2 * B])
        self.inp: Optional[Dict[str, torch.Tensor]] = None
        self._pf_stream = self._pf_event = self._pf_bufs = self._pf_batch = self._pf_picked = None
        self.inp_has_mask = self.fixed_mask = False
        self._sched: Dict[tuple, object] = {}
        if self.mask is not None:
            self._rand = torch.empty_like(self.mask.buf)
            self._keep = torch.empty_like(self.mask.buf, dtype=torch.bool)

    # ---- generator plan
    def _build_g(self):
        G, B, dev = self.G, self.B, self.device
        plan = Plan(dev, exact=self.exact, defer_bias=True, defer_pack=not self.exact, bf16=self.bf16)
        plan.direct_grads = self.g_bn
        self.plan = plan
        gp = G.global_pathway
        bufs = gp.alloc_concats(plan, B)
        self.bufs = bufs
        bufs["a128"].parts[2].requires_grad = False
        bufs["zin"].parts[1].requires_grad = False
        self.patches = [plan.new(B, h, w, 3, name=n, requires_grad=False) for n, (h, w) in zip(PART_NAMES, PATCH_HW)]
        self.patches_gt = [Act.empty(B, h, w, 3, dev) for (h, w) in PATCH_HW]
        self.frontal = Act.empty(B, 128, 128, 3, dev)
        self.t64, self.t32 = Act.empty(B, 64, 64, 3, dev), Act.empty(B, 32, 32, 3, dev)
        self.boxes = torch.zeros((B, 4, 4), dtype=torch.int32, device=dev)
        self.mask = None
        if self.use_dropout:
            self.mask = Act.empty(B, 1, 1, 256, dev)
        outs = G.trace_body(plan, bufs, self.patches, self.mask)
        self.fake, self.logits = outs[0], outs[1]
        self.local_imgs = list(outs[3:7])
        for t in [self.fake, self.logits] + self.local_imgs:
            plan.seed_grad(t)
        plan.trace_backward()

    def _ready_order(self) -> List[str]:
        """Parameter names in the order the backward plan finalises their gradients."""
        marks = self.plan.bwd_marks
        layers = sorted([L for L in self.plan.layers if L.name in marks], key=lambda L: marks[L.name])
        pname = {id(p): n for n, p in self.G.named_parameters()}
        order = []
        for L in layers:
            for p in (L.weight, L.bias):
                if p is not None and id(p) in pname:
                    order.append(pname[id(p)])
        return order

    # ---- one training step
    # ---- inputs: copied into static device buffers so that every pointer of the schedule is fixed
    INPUT_KEYS = ("img", "img_frontal", "img64_frontal", "img32_frontal", "landmarks", "z", "label", "gp_alpha")
    INPUT_KEYS_U8 = ("img_u8", "img_frontal_u8", "landmarks", "z", "label", "gp_alpha")

    def prefetch(self, b: Dict[str, torch.Tensor]):
        """Start the host->device copies of a FUTURE batch on a side stream (pinned host tensors), so they overlap the step
        that is running; a later step(b) with the same dict object picks the staged copies up with device-to-device copies."""
        if self._pf_stream is None:
            self._pf_stream = torch.cuda.Stream(device=self.device)
            self._pf_event = torch.cuda.Event()
            self._pf_bufs = {k: torch.empty_like(b[k], device=self.device).contiguous() for k in self.INPUT_KEYS}
        if self._pf_picked is not None:
            self._pf_stream.wait_event(self._pf_picked)   # the staging buffers are free once the last pick-up has run
        with torch.cuda.stream(self._pf_stream):
            for k in self.INPUT_KEYS:
                self._pf_bufs[k].copy_(b[k], non_blocking=True)
            self._pf_event.record(self._pf_stream)
        self._pf_batch = b

    def load_inputs(self, b: Dict[str, torch.Tensor]):
        if self.inp is None:
            self.inp = {k: torch.empty_like(b[k], device=self.device).contiguous() for k in self.INPUT_KEYS}
            if self.mask is not None:
                self.inp["dropout_mask"] = torch.ones_like(self.mask.buf)
        if b is self._pf_batch:     # staged by prefetch(): wait for the side stream, then device-to-device
            torch.cuda.current_stream(self.device).wait_event(self._pf_event)
            for k in self.INPUT_KEYS:
                self.inp[k].copy_(self._pf_bufs[k], non_blocking=True)
            if self._pf_picked is None:
                self._pf_picked = torch.cuda.Event()
            self._pf_picked.record(torch.cuda.current_stream(self.device))
            self._pf_batch = None
        else:
            for k in self.INPUT_KEYS:
                if b[k] is not self.inp[k]:
                    self.inp[k].copy_(b[k], non_blocking=True)
        self.inp_has_mask = "dropout_mask" in b
        if self.inp_has_mask:
            self.inp["dropout_mask"].copy_(b["dropout_mask"].reshape(self.mask.buf.shape))

    def _stage(self):
        """Static input buffers -> NHWC staging + landmark crops (all kernel launches, no host dependence)."""
        B, b = self.B, self.inp
        rt = not self.exact and not self.bf16      # bf16 mode rounds once, when the twin is written
        img = self.bufs["a128"].parts[2].act
        if self.input_format == "uint8":
            ops.u8_to_nhwc(b["img_u8"], img, rt)
            ops.u8_to_nhwc(b["img_frontal_u8"], self.frontal, False)
            ops.pyramid(self.frontal, self.t64, self.t32)
        else:
            img.from_nchw(b["img"], round_tf32=rt)
            self.frontal.from_nchw(b["img_frontal"], round_tf32=False)
            self.t64.from_nchw(b["img64_frontal"])
            self.t32.from_nchw(b["img32_frontal"])
        self.bufs["zin"].parts[1].act.from_nchw(b["z"].reshape(B, -1, 1, 1), round_tf32=rt)
        ops.patch_crop(img, b["landmarks"], [p.act for p in self.patches], self.boxes)
        ops.patch_crop(self.frontal, b["landmarks"], self.patches_gt, None)
        if self.mask is not None:
            if self.fixed_mask:
                self.mask.buf.copy_(b["dropout_mask"])
            else:
                p = self.G.feature_predict.dropout.p
                torch.rand(self.mask.buf.shape, out=self._rand)
                torch.ge(self._rand, p, out=self._keep)
                self.mask.buf.copy_(self._keep)
                self.mask.buf.mul_(1.0 / (1.0 - p))     # nn.Dropout's keep mask scaled by 1/(1-p); RNG = torch's philox

    def stage_inputs(self, b: Dict[str, torch.Tensor]):
        self.load_inputs(b)
        self.fixed_mask = self.inp_has_mask
        self._stage()

    def _schedule(self, optimize: bool) -> List[Callable]:
        """The whole step as a flat list of launches (Eager entries = collectives)."""
        B, w, crit = self.B, self.w, self.critic
        fake = self.fake.act
        x0 = crit.x0
        sch: List[Callable] = [self._stage]
        sch += self.plan.fwd
        # ---------------- D phase
        sch += [lambda: ops.view_copy(fake, _sl(x0, 0, B)), lambda: ops.view_copy(self.frontal, _sl(x0, B, 2 * B)),
                lambda: ops.lerp(self.frontal, fake, self.inp["gp_alpha"], _sl(x0, 2 * B, 3 * B))]
        sch += crit.d_phase_list(float(w["weight_gradient_penalty"]))
        sch.append(self.d_set.export)
        d_logits = self._d_logits
        sch.append(lambda: ops.view_copy(_sl(crit.logits, 0, 2 * B), Act(d_logits, 0, 1)))
        if self.world_size > 1 or self.force_reducer:
            sch.append(Collective(lambda: self._allreduce(self.flat_d.grad)))
        if optimize:
            sch.append(lambda: self.flat_d.adam(self.lr, 1.0 / self.world_size))
            sch.append(self.d_set.repack)
        # ---------------- G phase
        n128, n64, n32 = B * 3 * 128 * 128, B * 3 * 64 * 64, B * 3 * 32 * 32
        wp, ws, wt = w["weight_pixelwise"], w["weight_symmetry"], w["weight_total_varation"]
        coeffs = [wp * w["weight_128"] / n128, wp * w["weight_64"] / n64, wp * w["weight_32"] / n32,
                  ws * w["weight_128"] / n128, ws * w["weight_64"] / n64, ws * w["weight_32"] / n32,
                  wt / (B * 3 * 127 * 128), wt / (B * 3 * 128 * 127)]
        dfake = self.plan.grad_act(self.fake)
        sch.append(lambda: self.sums.zero_())
        sch.append(lambda: ops.image_losses(fake, self.frontal, self.t64, self.t32, dfake, coeffs, self.sums[0:8]))
        sch += crit.g_phase_list(float(w["weight_adv_G"]))
        if self.identity is not None:
            sch += self.identity.schedule()
        for i, (t, gt, (h, wd)) in enumerate(zip(self.local_imgs, self.patches_gt, PATCH_HW)):
            sch.append(lambda t=t, gt=gt, i=i, c=w["weight_pixelwise_local"] / (B * 3 * h * wd):
                       ops.l1_loss(t.act, gt, self.plan.grad_act(t), c, self.sums[8 + i:9 + i]))
        sch.append(lambda: ops.softmax_ce(self.logits.act, self.inp["label"], self.plan.grad_act(self.logits),
                                          w["weight_cross_entropy"] / B, self.sums[12:13]))
        sch.append(GradArena.get(self.device).zero)   # D's gradients were exported above
        if self.reducer is not None:   # bucketed all-reduce overlapped with backward
            hooks = self.reducer.hooks()
            for i, f in enumerate(self.plan.bwd):
                sch.append(f)
                if (i + 1) in hooks:
                    sch.append(Collective(hooks[i + 1]))
            sch.append(Collective(self.reducer.finish))
        else:
            sch += self.plan.bwd
            sch.append(self.g_set.export)
            if self.world_size > 1:     # one all-reduce of the whole flat gradient after backward
                sch.append(Collective(lambda: self._allreduce(self.flat_g.grad)))
        if optimize:
            sch.append(lambda: self.flat_g.adam(self.lr, 1.0 / self.world_size))
            sch.append(self.g_set.repack)
        return sch

    def step(self, b: Dict[str, torch.Tensor], optimize: bool = True, read_metrics: bool = True,
             prefetch_next: Optional[Dict[str, torch.Tensor]] = None):
        """One G+D training step.  With use_graphs the schedule is captured into CUDA graphs on first use and replayed.
        prefetch_next: the batch of the NEXT step (pinned host tensors); its host->device copies are started on a side
        stream as soon as this step's launches are enqueued (see prefetch())."""
        self.load_inputs(b)
        self.fixed_mask = self.inp_has_mask
        if self.identity is not None:
            self.identity.base.refresh_folded()   # re-fold the frozen network if its weights were replaced (load_state_dict)
        key = (optimize, self.fixed_mask)
        if key not in self._sched:
            sch = self._schedule(optimize)
            self._sched[key] = GraphRunner(sch, capture_collectives=self.graph_collectives) if self.use_graphs else sch
        sch = self._sched[key]
        if self.use_graphs:
            sch.run()
        else:
            for f in sch:
                f()
        if optimize or self.g_bn:
            self.steps += 1      # generator forwards in training mode (BatchNorm statistics updates), see sync_buffers()
        if prefetch_next is not None:
            self.prefetch(prefetch_next)
        return self.read_metrics() if read_metrics else None

    def sync_buffers(self):
        """num_batches_tracked of the generator's BatchNorm layers (state_dict parity with nn.BatchNorm2d in train mode; the
        running statistics themselves are updated in place by the forward kernels)."""
        for m in self.G.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None:
                m.num_batches_tracked.fill_(self.steps)

    def _allreduce(self, t: torch.Tensor):
        import torch.distributed as dist
        dist.all_reduce(t, group=self.group)

    # ---- metrics (small device->host reads; all arithmetic on the host)
    def read_metrics(self) -> Dict[str, float]:
        B, w, crit = self.B, self.w, self.critic
        s = self.sums.cpu().tolist()
        n128, n64, n32 = B * 3 * 128 * 128, B * 3 * 64 * 64, B * 3 * 32 * 32
        pixel = w["weight_128"] * s[0] / n128 + w["weight_64"] * s[1] / n64 + w["weight_32"] * s[2] / n32
        sym = w["weight_128"] * s[3] / n128 + w["weight_64"] * s[4] / n64 + w["weight_32"] * s[5] / n32
        tv = s[6] / (B * 3 * 127 * 128) + s[7] / (B * 3 * 128 * 127)
        local = sum(s[8 + i] / (B * 3 * h * wd) for i, (h, wd) in enumerate(PATCH_HW))
        ce = s[12] / B
        dl = self._d_logits.cpu()[..., 0].reshape(2 * B, -1).mean(1).tolist()       # host arithmetic on 2B*16 floats
        d_fake, d_real = sum(dl[:B]) / B, sum(dl[B:]) / B
        gp = float(crit.gp_sum.cpu()[0]) / B
        adv_g = -float(crit.logits.buf[:B].cpu()[..., 0].mean())
        g_total = w["weight_pixelwise"] * pixel + w["weight_pixelwise_local"] * local + w["weight_symmetry"] * sym + \
            w["weight_adv_G"] * adv_g + w["weight_total_varation"] * tv + w["weight_cross_entropy"] * ce
        out = dict(pixel=pixel, local=local, symmetry=sym, tv=tv, ce=ce, adv_g=adv_g, g_total=g_total, d_fake=d_fake,
                   d_real=d_real, gp=gp, d_total=d_fake - d_real + w["weight_gradient_penalty"] * gp)
        if self.identity is not None:
            out["ip"] = self.identity.value(s[13:15])
            out["g_total"] = g_total + w["weight_identity_preserving"] * out["ip"]
        return out
