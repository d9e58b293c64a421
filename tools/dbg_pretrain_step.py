"""Developer tool: per-parameter gradient / update deviation of PretrainTrainer (fp32-exact mode) from the oracle step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import pretrain_port as P
from tpgan_b200.MobileNetV2 import MobileNetV2
from tpgan_b200.pretrain_step import PretrainTrainer

exact = "tf32" not in sys.argv
torch.manual_seed(4)
port = P.MobileNetV2Port()
net = MobileNetV2()
net.load_state_dict(port.state_dict())
net.cuda()
B = 4
tr = PretrainTrainer(net, B, exact=exact)
opt = torch.optim.SGD(port.parameters(), **P.SGD)
p0 = {k: v.detach().clone() for k, v in port.named_parameters()}
x, true, u = P.make_batch(B, seed=20)
want, labs, _, _ = P.pretrain_step(port, x, true, u, None)
m = tr.step(x.cuda(), true.cuda(), u.cuda(), optimize=False)
print("loss", m, float(want), "labels agree", float((tr.labels.cpu() == torch.stack(labs)).float().mean()))
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
rows = []
pg = dict(net.named_parameters())
for k, p in port.named_parameters():
    rows.append((rel(pg[k].grad, p.grad), k, float(p.grad.norm()), float(pg[k].grad.norm())))
rows.sort(reverse=True)
for r in rows[:25]:
    print("%.3e %-45s |ref| %.3e |got| %.3e" % r)
print("median", sorted(r[0] for r in rows)[len(rows) // 2])
# one optimisation step
opt.step()
tr.step(x.cuda(), true.cuda(), u.cuda(), optimize=True)
rows = []
for k, p in port.named_parameters():
    rows.append((rel(pg[k].detach() - p0[k].cuda(), p.detach() - p0[k]), k))
rows.sort(reverse=True)
print("update deviations:", rows[:8], "median", rows[len(rows) // 2])
