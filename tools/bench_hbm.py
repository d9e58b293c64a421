"""Achieved HBM bandwidth of the memory-bound kernels of the path: landmark patch crop, LocalFuser stitch
forward / backward, fused image losses, bias gradient, Adam - algorithmic bytes (the figures of DESIGN.md section 4) over
the CUDA-event time per launch, next to the measured copy bandwidth of MEASURED_PEAKS.json.
Usage: python tools/bench_hbm.py [--batch 32] [--out file.jsonl]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpgan_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)   # > L2-sized write between timed launches
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
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []

    def report(name, nbytes, fn):
        ms = timeit(fn, a.iters, flush)
        r = dict(kernel=name, batch=B, ms=round(ms, 4), algorithmic_mb=round(nbytes / 1e6, 2),
                 gbs=round(nbytes / ms / 1e6, 1), frac_hbm_peak=round(nbytes / ms / 1e6 / peak, 3))
        rows.append(r)
        print(json.dumps(r), flush=True)

    HW = [(40, 40), (40, 40), (32, 40), (32, 48)]
    npatch = sum(h * w for h, w in HW)          # 6016 pixels
    # ---- LocalFuser stitch, C = 64 (reads the four patches, writes the 128x128 map + 1-byte arg-max)
    for C in (64, 3):
        patches = [ops.Act.empty(B, h, w, C) for h, w in HW]
        for q in patches:
            q.buf.uniform_(-1, 1, generator=g)
        out = ops.Act.empty(B, 128, 128, C)
        arg = torch.empty((B, 128, 128, C), dtype=torch.uint8, device="cuda")
        report(f"local_fuse fwd C={C}", B * (npatch + 16384) * C * 4 + B * 16384 * C, lambda: ops.local_fuse(patches, out, arg))
        dout = ops.Act.empty(B, 128, 128, C)
        dout.buf.uniform_(-1, 1, generator=g)
        dp = [ops.Act.empty(B, h, w, C) for h, w in HW]
        report(f"local_fuse bwd C={C}", B * npatch * C * (4 + 4 + 1), lambda: ops.local_fuse_backward(dout, arg, dp))
    # ---- landmark patch crop (reads / writes the 6016 patch pixels of a 3-channel image)
    img = ops.Act.empty(B, 128, 128, 3)
    img.buf.uniform_(-1, 1, generator=g)
    lm = torch.tensor([[39.5, 40.3], [86.2, 39.9], [63.4, 63.0], [44.1, 87.2], [82.0, 88.1]], device="cuda").repeat(B, 1, 1).contiguous()
    pt = [ops.Act.empty(B, h, w, 3) for h, w in HW]
    boxes = torch.zeros((B, 4, 4), dtype=torch.int32, device="cuda")
    report("patch_crop", 2 * B * npatch * 3 * 4, lambda: ops.patch_crop(img, lm, pt, boxes))
    # ---- fused image losses (fake + three targets read once, d fake written)
    t128, t64, t32 = ops.Act.empty(B, 128, 128, 3), ops.Act.empty(B, 64, 64, 3), ops.Act.empty(B, 32, 32, 3)
    for t in (t128, t64, t32):
        t.buf.uniform_(-1, 1, generator=g)
    dfake = ops.Act.empty(B, 128, 128, 3)
    sums = torch.zeros(8, device="cuda")
    nb = B * 3 * 4 * (128 * 128 * 3 + 64 * 64 + 32 * 32)
    report("image_losses", nb, lambda: ops.image_losses(img, t128, t64, t32, dfake, [1e-6] * 8, sums))
    # ---- Adam over a 138 M parameter flat buffer (28 B / parameter)
    n = 137_764_238
    pbuf = [torch.zeros(n, device="cuda") for _ in range(4)]
    step = torch.ones(1, dtype=torch.int32, device="cuda")
    report("adam (G, 137.8 M params)", 28 * n, lambda: ops.adam_step_dev(pbuf[0], pbuf[1], pbuf[2], pbuf[3], 1e-4, 0.9, 0.999, 1e-8, 0.0, step))
    del pbuf
    # ---- bias gradient of a 64-channel 128x128 activation gradient
    dy = ops.Act.empty(B, 128, 128, 64)
    dy.buf.uniform_(-1, 1, generator=g)
    db = torch.zeros(64, device="cuda")
    tab = ops.JobTable("bias", [ops.bias_job(dy, db)], torch.device("cuda"))
    report("bias_grad_multi 64ch@128x128", B * 16384 * 64 * 4, tab.run)
    assert _lib.kernel_status() == 0
    if a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
