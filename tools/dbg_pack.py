import sys, torch
sys.path.insert(0, '/root/repo')
from tpgan_b200 import ops
w = torch.randn(512, 512, 8, 8, device='cuda')
pk = ops.pack_weights(w, ops.CONV_FWD)
g = torch.zeros_like(w)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print("pack fc1 ms", t(lambda: ops.pack_weights(w, ops.CONV_FWD, pk)))
print("unpack fc1 ms", t(lambda: ops.unpack_weights(pk, g, ops.CONV_FWD)))
ops.unpack_weights(pk, g, ops.CONV_FWD)
from oracle.model_port import tf32_rna
print("roundtrip exact", bool(torch.equal(g.cpu(), tf32_rna(w.cpu()))))
