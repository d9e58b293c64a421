"""Developer tool: per-layer deviation of the traced MobileNetV2 plan from the oracle port (forward, training mode)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tpgan_b200.D_and_G_model as M
from oracle.pretrain_port import MobileNetV2Port, make_batch
from tpgan_b200.engine import Plan
from tpgan_b200.MobileNetV2 import MobileNetV2

exact = len(sys.argv) > 1 and sys.argv[1] == "exact"
B = 4
torch.manual_seed(0)
port = MobileNetV2Port()
net = MobileNetV2()
net.load_state_dict(port.state_dict())
net.cuda().train()
port.train()
x, _, _ = make_batch(B, seed=7)
acts = {}
for name, mod in port.named_modules():
    if isinstance(mod, (torch.nn.Conv2d, torch.nn.BatchNorm2d, torch.nn.ReLU6)):
        mod.register_forward_hook(lambda m, i, o, name=name: acts.__setitem__(name, o.detach().clone()))
with torch.no_grad():
    lw, cw = port(x)
plan = Plan(torch.device("cuda"), training=True, need_wgrad=False, exact=exact)
t = plan.new(B, 128, 128, 3, name="in", requires_grad=False)
loc, cls = net.trace(plan, t)
t.act.from_nchw(x.cuda(), round_tf32=not exact)
plan.run_forward()
torch.cuda.synchronize()


def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


for name, T in plan.named.items():
    ref = acts.get(name)
    if ref is None:
        continue
    # a BN layer's plan output includes the ReLU6 / residual: compare with the next module's output where there is one
    got = T.act.to_nchw()
    parts = name.split(".")
    nxt = ".".join(parts[:-1] + [str(int(parts[-1]) + 1)]) if parts[-1].isdigit() else None
    if isinstance(dict(port.named_modules())[name], torch.nn.BatchNorm2d) and nxt in acts:
        ref = acts[nxt]
    print(f"{name:40s} {tuple(got.shape)!s:22s} rel {rel(got, ref):.3e}")
n = loc.act.c // 2
print("loc", rel(loc.act.buf.view(B, -1)[:, :2 * n].reshape(B, n, 2), lw), "cls",
      rel(cls.act.buf.view(B, -1)[:, :5 * n].reshape(B, n, 5), cw))
