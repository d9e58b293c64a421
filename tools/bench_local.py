"""Micro-benchmark of the grouped local-pathway launches (four problems per launch: eyes 40x40, nose 32x40, mouth 32x48)
through the C ABI: forward, input gradient and weight gradient of one layer shape.
Usage: python tools/bench_local.py [--cin 64 --cout 64 --k 3 --div 1] (div = spatial down-scale 1/2/4/8 of the patch)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpgan_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--div", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kinds", default="fwd,dgrad,wgrad")
    ap.add_argument("--chain", type=int, default=0, help="also time N back-to-back launches replayed as one CUDA graph (warm caches)")
    a = ap.parse_args()
    shapes = [(40 // a.div, 40 // a.div), (40 // a.div, 40 // a.div), (32 // a.div, 40 // a.div), (32 // a.div, 48 // a.div)]
    B, k, p = a.batch, a.k, (a.k - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(0)
    keep, fa, da, wa = [], [], [], []
    macs = 0
    for (h, w) in shapes:
        x = ops.Act.empty(B, h, w, a.cin)
        x.buf.uniform_(-1, 1, generator=g)
        y, dy, dx = ops.Act.empty(B, h, w, a.cout), ops.Act.empty(B, h, w, a.cout), ops.Act.empty(B, h, w, a.cin)
        dy.buf.uniform_(-1, 1, generator=g)
        wt = torch.empty((a.cout, a.cin, k, k), device="cuda").uniform_(-0.05, 0.05, generator=g)
        bias = torch.zeros(ops.round_up(a.cout, 4), device="cuda")
        wf, wd = ops.pack_weights(wt, ops.CONV_FWD), ops.pack_weights(wt, ops.CONV_DGRAD)
        dw = ops.alloc_packed(ops.CONV_FWD, tuple(wt.shape))
        keep += [x, y, dy, dx, wt, bias, wf, wd, dw]
        fa.append(ops.conv_args(ops.CONV_FWD, x, y, wf, k, 1, p, bias=bias, slope=0.01, epilogue=ops.EPI_LEAKY))
        da.append(ops.conv_args(ops.CONV_DGRAD, dy, dx, wd, k, 1, p, mask=x, slope=0.01, epilogue=ops.EPI_MASK))
        wa.append(ops.wgrad_args(ops.CONV_FWD, x, dy, dw, k, 1, p))
        macs += B * h * w * a.cin * a.cout * k * k
    fns = {"fwd": lambda: ops.conv2d_grouped(fa), "dgrad": lambda: ops.conv2d_grouped(da), "wgrad": lambda: ops.wgrad_grouped(wa)}
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")

    def graph_ms(fn, reps=8):
        # device-side time of one launch: (flush + launch) x reps replayed as a CUDA graph minus the flushes alone - the eager
        # event timing below includes the host's planning of a 4-group launch (tens of microseconds), the step does not
        def cap(body):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    flush.add_(1.0)
                    body()
            return g
        gs = [cap(fn), cap(lambda: None)]
        out = []
        for g in gs:
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            out.append(min(ts))
        return (out[0] - out[1]) / reps

    for kind in a.kinds.split(","):
        fn = fns[kind]
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        gms = graph_ms(fn)
        cms = None
        if a.chain:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(a.chain):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            cms = e0.elapsed_time(e1) / a.chain
        ts = []
        for _ in range(a.iters):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(json.dumps(dict(kind=kind, cin=a.cin, cout=a.cout, k=k, div=a.div, kernel=_lib.last_conv_kernel() if kind != "wgrad" else "wgrad",
                              ms=round(ms, 4), tflops=round(2 * macs / ms / 1e9, 1), graph_ms=round(gms, 4),
                              graph_tflops=round(2 * macs / gms / 1e9, 1), chain_ms=None if cms is None else round(cms, 4))), flush=True)
    assert _lib.kernel_status() == 0


if __name__ == "__main__":
    main()
