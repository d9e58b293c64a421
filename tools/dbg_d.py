import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import step as ostep, model_port as mp
from tpgan_b200 import D_and_G_model as M, config, _lib
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
mp.EMULATE_TF32 = True
torch.manual_seed(0)
D = M.Discriminator(False)
sd = {k: v.clone() for k, v in D.state_dict().items()}
b = ostep.make_batch(2)
D.cuda()
x = b["img"]
with torch.no_grad():
    out = D(x.cuda())
    plan = list(D._cache().plans.values())[0].plan
    # port, layer by layer
    h = mp._q(x)
    acts = {}
    for i in range(4):
        h = mp._conv(sd, f"model.{i}", h, 2, 1); acts[f"model.{i}"] = h
    h0 = h
    hh = mp._conv(sd, "model.4.layers.0", h, 1, 1); acts["model.4.layers.0"] = hh
    h = mp._res(sd, "model.4", h, 3); acts["model.4.layers.1"] = h
    h = mp._conv(sd, "model.5", h, 2, 1); acts["model.5"] = h
    hh = mp._conv(sd, "model.6.layers.0", h, 1, 1); acts["model.6.layers.0"] = hh
    h = mp._res(sd, "model.6", h, 3); acts["model.6.layers.1"] = h
    h = mp._conv(sd, "model.7", h, 1, 1, None); acts["model.7"] = h
    for k, v in acts.items():
        t = plan.named[k]
        g = t.act.to_nchw().cpu()
        d = (g - v).abs()
        print(k, tuple(v.shape), "rel %.2e" % rel(g, v), "max abs %.3e" % float(d.max()), "ref absmax %.2f" % float(v.abs().max()))
