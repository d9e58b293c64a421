"""Per-shape micro-benchmark of the tensor-core convolution kernels (fwd / dgrad / wgrad) through the C ABI.

Shapes are SURVEY.md table T1 (the layers that carry 95 % of the FLOPs of the G+D step).  Reports, per kernel, the CUDA-event
time per launch and the algorithmic TFLOP/s (2*M*N*K with the unpadded logical dims) next to the tf32 tensor peak
(half of the measured bf16 figure in MEASURED_PEAKS.json).  Usage:  python tools/bench_conv.py [--batch 32] [--only NAME]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpgan_b200 import ops  # noqa: E402

# name, cin, cout, k, stride, pad, H(=W), transposed
SHAPES = [
    ("enh128_206x206_k5", 206, 206, 5, 1, 2, 128, False),
    ("add128_75x75_k7", 75, 75, 7, 1, 3, 128, False),
    ("conv0rb_64x64_k7", 64, 64, 7, 1, 3, 128, False),
    ("enh32_416x416_k3", 416, 416, 3, 1, 1, 32, False),
    ("enh64_208x208_k3", 208, 208, 3, 1, 1, 64, False),
    ("enh16_768x768_k3", 768, 768, 3, 1, 1, 16, False),
    ("conv5_206x64_k5", 206, 64, 5, 1, 2, 128, False),
    ("conv00_3x64_k7", 3, 64, 7, 1, 3, 128, False),
    ("conv4rb_512x512_k3", 512, 512, 3, 1, 1, 8, False),
    ("add64_80x80_k5", 80, 80, 5, 1, 2, 64, False),
    ("conv5rb_64x64_k3", 64, 64, 3, 1, 1, 128, False),
    ("conv1_64x64_k5s2", 64, 64, 5, 2, 2, 128, False),
    ("up128_208x64_d3s2", 208, 64, 3, 2, 1, 64, True),
    ("up32_768x256_d3s2", 768, 256, 3, 2, 1, 16, True),
]


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)  # > L2-sized write between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--kinds", default="fwd,dgrad,wgrad")
    ap.add_argument("--out", default="")
    ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"])
    a = ap.parse_args()
    bf16 = a.dtype == "bf16"
    arena = ops.Arena("cuda", shadow=bf16)
    ops._ARENA.append(arena)      # every buffer below comes from the (shadowed, in bf16 mode) arena
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    tf32_peak = peaks.get("bf16_tflops", 1590.0) / (1 if bf16 else 2)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")  # 256 MB
    B = a.batch
    rows = []
    for name, cin, cout, k, s, pad, H, tr in SHAPES:
        if a.only and a.only not in name:
            continue
        if tr:
            Ho = H * s
            wshape = (cin, cout, k, k)
            kf, kd = ops.DECONV_FWD, ops.DECONV_DGRAD
            macs = B * H * H * cin * cout * k * k
        else:
            Ho = (H + 2 * pad - k) // s + 1
            wshape = (cout, cin, k, k)
            kf, kd = ops.CONV_FWD, ops.CONV_DGRAD
            macs = B * Ho * Ho * cin * cout * k * k
        g = torch.Generator(device="cuda").manual_seed(0)
        x = ops.Act.empty(B, H, H, cin)
        x.buf.uniform_(-1, 1, generator=g)
        if cin % 4:
            x.buf[..., cin:] = 0
        y = ops.Act.empty(B, Ho, Ho, cout)
        dy = ops.Act.empty(B, Ho, Ho, cout)
        dy.buf.uniform_(-1, 1, generator=g)
        if cout % 4:
            dy.buf[..., cout:] = 0
        dx = ops.Act.empty(B, H, H, cin)
        w = torch.empty(wshape, device="cuda").uniform_(-0.05, 0.05, generator=g)
        bias = torch.zeros(ops.round_up(cout, 4), device="cuda")
        wf = ops.pack_weights(w, kf, round_tf32=not bf16)
        wd = ops.pack_weights(w, kd, round_tf32=not bf16)
        if bf16:
            wf16, wd16 = ops.alloc_packed16(wf), ops.alloc_packed16(wd)
            ops.cast_packed(wf, wf16)
            ops.cast_packed(wd, wd16)
            ops.cast_bf16(x)
            ops.cast_bf16(dy)
            wf, wd = wf16, wd16
        dw = ops.alloc_packed(kf, wshape)
        o32 = os.environ.get("BENCH_NO_FP32_OUT") is None
        fns = {
            "fwd": lambda: ops.conv2d(kf, x, y, wf, k, s, pad, bias=bias, slope=0.01, epilogue=ops.EPI_LEAKY,
                                      **(dict(bf16=True, out32=o32) if bf16 else {})),
            "dgrad": lambda: ops.conv2d(kd, dy, dx, wd, k, s, pad, mask=x, slope=0.01, epilogue=ops.EPI_MASK,
                                        **(dict(bf16=True, out32=o32) if bf16 else {})),
            "wgrad": lambda: ops.wgrad(kf, x, dy, dw, k, s, pad, bf16=bf16),
        }
        for kind in a.kinds.split(","):
            ms = timeit(fns[kind], a.iters, flush)
            tf = 2 * macs / ms / 1e9
            row = dict(shape=name, kind=kind, dtype=a.dtype, batch=B, ms=round(ms, 4), tflops=round(tf, 1), frac_peak=round(tf / tf32_peak, 3))
            rows.append(row)
            print(json.dumps(row), flush=True)
        del x, y, dy, dx, w, wf, wd, dw
        arena.chunk = None          # start a fresh chunk per shape: the previous one is freed with its last view
        arena.twins.clear()
        torch.cuda.empty_cache()
    from tpgan_b200 import _lib
    assert _lib.kernel_status() == 0
    if a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
