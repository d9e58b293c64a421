"""Achieved HBM bandwidth of the memory-bound kernels of the Pretrain path (row a14) on the largest layer shapes of
MobileNetV2 at 128x128: training BatchNorm forward / backward, depthwise 3x3 conv forward / dgrad / wgrad, SGD-Nesterov,
MultiTaskLoss - algorithmic bytes over the CUDA-event time per call (each call = the launches the C-ABI entry point makes),
L2 flushed between calls.  Also the ncu target for the `--set full` captures of these kernels.
Usage: python tools/bench_pretrain_kernels.py [--batch 32] [--out file.jsonl]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpgan_b200 import ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)   # > L2-sized write between timed calls
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
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    B = a.batch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 6445.0
    p = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    rows = []

    def report(name, nbytes, fn):
        ms = timeit(fn, a.iters, flush)
        r = dict(kernel=name, batch=B, ms=round(ms, 4), algorithmic_mb=round(nbytes / 1e6, 2),
                 gbs=round(nbytes / ms / 1e6, 1), frac_hbm_peak=round(nbytes / ms / 1e6 / peak, 3))
        rows.append(r)
        print(json.dumps(r), flush=True)

    rnd = lambda *s: torch.randn(*s, device="cuda")
    # (C, H, W): expand output of bottleneck 1 (the largest tensor of the network), a mid layer, a late layer
    for C, H, W in ((96, 64, 64), (144, 32, 32), (384, 8, 8)):
        x = ops.Act(rnd(B, H, W, C))
        y, dy, dx = x.like(), ops.Act(rnd(B, H, W, C)), x.like()
        gamma, beta = torch.rand(C, device="cuda") + 0.5, rnd(C)
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        sums = torch.zeros(2 * C + 1, dtype=torch.float64, device="cuda")
        dsums = torch.zeros_like(sums)
        coef = torch.zeros(4 * C, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        el = B * H * W * C * 4.0
        report(f"bn_forward train+relu6 {C}x{H}x{W} (read x twice, write y)", 3 * el,
               lambda: ops.bn_forward(x, None, y, gamma, beta, rm, rv, 0.1, 1e-5, True, True, True, sums, coef))
        report(f"bn_backward train+relu6 {C}x{H}x{W} (read dy, x twice, write dx)", 5 * el,
               lambda: ops.bn_backward(dy, x, dx, coef, True, True, False, True, dsums, dg, db))
        w = rnd(C, 1, 3, 3)
        dw = torch.zeros_like(w)
        for s in (1, 2):
            Ho, Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
            yo, dyo = ops.Act.empty(B, Ho, Wo, C), ops.Act(rnd(B, Ho, Wo, C))
            eo = B * Ho * Wo * C * 4.0
            report(f"dwconv3x3 fwd s{s} {C}x{H}x{W}", el + eo, lambda: ops.dwconv3x3(x, yo, w, s))
            report(f"dwconv3x3 dgrad s{s} {C}x{H}x{W}", el + eo, lambda: ops.dwconv3x3_dgrad(dyo, dx, w, s, False))
            report(f"dwconv3x3 wgrad s{s} {C}x{H}x{W}", el + eo, lambda: ops.dwconv3x3_wgrad(x, dyo, dw, s))
    n = 7_676_344
    pbuf, g, m = rnd(n), rnd(n), torch.zeros(n, device="cuda")
    lr = torch.full((1,), 5e-4, device="cuda")
    report("sgd_step nesterov (7.68 M params: read p, g, buf; write p, buf)", 20.0 * n,
           lambda: ops.sgd_step(pbuf, g, m, lr, 0.9, 5e-4, True))
    npt = 394
    loc, cls = torch.rand(B, npt, 2, device="cuda") * 128, rnd(B, npt, 5)
    true, u = torch.rand(B, 8, device="cuda") * 128, torch.rand(B, npt, device="cuda")
    dloc, dcls = torch.empty_like(loc), torch.empty_like(cls)
    labels = torch.zeros(B, npt, dtype=torch.int32, device="cuda")
    sums3 = torch.zeros(3, device="cuda")
    report("multitask_loss (one CTA per sample; latency-bound: O(n^2) rank counting in shared memory)",
           B * npt * (2 + 5 + 1 + 2 + 5 + 1) * 4.0,
           lambda: ops.multitask_loss(loc, cls, true, u, npt, 2 * npt, 5 * npt, 39, 128.0, 128.0, 30.0, 0.1, 5.0, 1.0 / B,
                                      dloc, dcls, labels, sums3))
    if a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
