import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tpgan_b200 import ops
from oracle.model_port import tf32_rna
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
g = torch.Generator().manual_seed(0)
for (n, cin, cout, h, k, s) in [(2, 64, 64, 32, 3, 1), (2, 3, 64, 128, 3, 2), (2, 64, 128, 64, 3, 2), (2, 3, 64, 128, 7, 1), (2, 512, 512, 8, 3, 1)]:
    x0 = torch.rand((n, cin, h, h), generator=g) * 2 - 1
    w0 = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) / (cin * k * k) ** 0.5
    b = torch.rand(cout, generator=g) - 0.5
    x, w = tf32_rna(x0), tf32_rna(w0)
    ref = tf32_rna(F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=k // 2).float(), 0.01))
    xa = ops.Act.empty(n, h, h, cin).from_nchw(x0.cuda(), round_tf32=True)
    print("  input rounding equal:", torch.equal(xa.to_nchw().cpu(), x))
    ho = ref.shape[2]
    out = ops.Act.empty(n, ho, ho, cout)
    pw = ops.pack_weights(w0.cuda(), ops.CONV_FWD, round_tf32=True)
    bb = torch.zeros(ops.round_up(cout, 4)); bb[:cout] = b
    ops.conv2d(ops.CONV_FWD, xa, out, pw, k, s, k // 2, bias=bb.cuda(), slope=0.01, epilogue=ops.EPI_LEAKY, round_tf32=True)
    torch.cuda.synchronize()
    o = out.to_nchw().cpu()
    print((n, cin, cout, h, k, s), "gpu vs emulated", rel(o, ref), "max abs", float((o - ref).abs().max()))
