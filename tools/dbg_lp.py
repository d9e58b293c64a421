import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import step as ostep, model_port as mp
from tpgan_b200 import D_and_G_model as M, config, _lib
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
M.EXACT_MODE = True
torch.manual_seed(0)
P = M.LocalPathway(False)
sd = {"lp." + k: v.clone() for k, v in P.state_dict().items()}
b = ostep.make_batch(2)
P.cuda()
x = b["nose"]
with torch.no_grad():
    img, feat = P(x.cuda())
    plan = list(P._cache().plans.values())[0].plan
    rimg, rfeat = mp.local_pathway(sd, "lp", x)
    print("img", rel(img, rimg), "feat", rel(feat, rfeat))
    # layer by layer
    c = lambda name, t, s: mp._res(sd, f"lp.{name}.1", mp._conv(sd, f"lp.{name}.0", t, s, 1), 3)
    a0 = mp._conv(sd, "lp.conv0.0", x, 1, 1)
    print("conv0.0", rel(plan.named["local_pathway.conv0.0"].act.to_nchw(), a0))
    h = mp._conv(sd, "lp.conv0.1.layers.0", a0, 1, 1)
    print("conv0.1.layers.0", rel(plan.named["local_pathway.conv0.1.layers.0"].act.to_nchw(), h))
    conv0 = c("conv0", x, 1)
    print("conv0", rel(plan.named["local_pathway.conv0.1.layers.1"].act.to_nchw(), conv0))
    conv1 = c("conv1", conv0, 2)
    print("conv1", rel(plan.named["local_pathway.conv1.1.layers.1"].act.to_nchw(), conv1))
    conv2 = c("conv2", conv1, 2)
    print("conv2", rel(plan.named["local_pathway.conv2.1.layers.1"].act.to_nchw(), conv2))
    conv3 = c("conv3", conv2, 2)
    print("conv3", rel(plan.named["local_pathway.conv3.1.layers.1"].act.to_nchw(), conv3))
    d0 = mp._deconv(sd, "lp.deconv0", conv3, 2, 1, 1)
    print("deconv0", rel(plan.named["local_pathway.deconv0"].act.to_nchw(), d0))
print("status", _lib.kernel_status())
