import sys, torch
sys.path.insert(0, '.')
from tpgan_b200 import ops, _lib
torch.manual_seed(0)
n, cin, cout, h, w, k, s, p = 2, 32, 16, 8, 8, 1, 1, 0
x = torch.rand(n, cin, h, w) * 2 - 1
dy = torch.rand(n, cout, h, w) * 2 - 1
ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, stride=s, padding=p)
xa = ops.Act.empty(n, h, w, cin).from_nchw(x.cuda())
dya = ops.Act.empty(n, h, w, cout).from_nchw(dy.cuda())
dw = ops.alloc_packed(ops.CONV_FWD, (cout, cin, k, k))
print("packed shape", dw.data.shape, dw.rows_pad, dw.k_pad)
ops.wgrad(ops.CONV_FWD, xa, dya, dw, k, s, p)
torch.cuda.synchronize()
print("status", _lib.kernel_status())
d = dw.data.cpu()
print("dw abs sum per tap", d.abs().sum(dim=(1, 2)))
print("dw[0,:4,:6]\n", d[0, :4, :6])
print("ref[:4,:6]\n", ref[:4, :6, 0, 0])
got = torch.zeros((cout, cin, k, k), device="cuda")
ops.unpack_weights(dw, got, ops.CONV_FWD)
torch.cuda.synchronize()
print("got[:4,:6]\n", got[:4, :6, 0, 0].cpu())
# pack/unpack roundtrip
wt = torch.rand(cout, cin, 3, 3).cuda()
pk = ops.pack_weights(wt, ops.CONV_FWD, round_tf32=False)
back = torch.zeros_like(wt)
ops.unpack_weights(pk, back, ops.CONV_FWD)
torch.cuda.synchronize()
print("roundtrip max err", float((back - wt).abs().max()))
