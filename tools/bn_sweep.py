"""Developer tool: time the BatchNorm statistics / backward-reduce kernels alone on the largest layer for a grid-size sweep
(TPGAN_BN_PER_SM set by the caller)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpgan_b200 import ops
B, C, H, W = 32, 96, 64, 64
x = ops.Act(torch.randn(B, H, W, C, device="cuda")); y = x.like(); dy = ops.Act(torch.randn(B, H, W, C, device="cuda")); dx = x.like()
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
sums = torch.zeros(2 * C + 1, dtype=torch.float64, device="cuda"); dsums = torch.zeros_like(sums)
coef = torch.zeros(4 * C, device="cuda"); dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
flush = torch.zeros(64 << 20, device="cuda")
def t(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(9):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[4]
f = t(lambda: ops.bn_forward(x, None, y, gamma, beta, rm, rv, 0.1, 1e-5, True, True, True, sums, coef))
b = t(lambda: ops.bn_backward(dy, x, dx, coef, True, True, False, True, dsums, dg, db))
print(os.environ.get("TPGAN_BN_PER_SM"), "fwd(stats+apply) us", round(f * 1e3, 1), "bwd(reduce+apply) us", round(b * 1e3, 1))
