"""Debug driver: traced Generator / Discriminator vs the fp32 CPU port (oracle/model_port.py) on a seeded batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import step as ostep, model_port as mp
from tpgan_b200 import D_and_G_model as M, config, _lib

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mp.EMULATE_TF32 = (os.environ.get('EMU', '0') == '1')
M.EXACT_MODE = (os.environ.get('EXACT', '1') == '1')
torch.manual_seed(0)
G = M.Generator(config.G['zdim'], config.G['num_classes'], config.G['use_batchnorm'], config.G['use_residual_block'])
D = M.Discriminator(config.D['use_batchnorm'])
sd_g = {k: v.clone() for k, v in G.state_dict().items()}
sd_d = {k: v.clone() for k, v in D.state_dict().items()}
b = ostep.make_batch(B)
G.cuda(); D.cuda()
names = ("img", "left_eye", "right_eye", "nose", "mouth", "z")
cu = [b[k].cuda() for k in names]
t0 = time.time()
outs = G(*cu, False)
torch.cuda.synchronize()
print("G fwd ok", time.time() - t0, "kernel_status", _lib.kernel_status())
# CPU port with autograd; activation backward uses the sign pattern of the CUDA path's stored activations
plan = list(G._cache().plans.values())[0].plan
def hook(name):
    t = plan.named.get(name)
    if t is None:
        print("  (no mask for", name, ")")
        return None
    a = t.act
    if name.endswith("deconv_8"):
        a = M.Act(a.buf.view(a.n, 8, 8, 64))
    o = a.to_nchw().cpu()
    if t.cmap is not None:
        idx = [i for i, c in enumerate(t.cmap) if c >= 0]
        o = o[:, idx]
    return o
if os.environ.get('MASK', '1') == '1':
    mp.MASK_HOOK = hook
pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
ref = mp.generator(pg, *[b[k] for k in names])
onames = ("fake", "logits", "fused_fake", "le", "re", "nose", "mouth", "fused_in")
for n, o, r in zip(onames, outs, ref):
    print(f"  {n:10s} {tuple(o.shape)} rel {rel(o, r):.2e}")
# backward with fixed cotangents
gen = torch.Generator().manual_seed(7)
cots = [torch.randn(r.shape, generator=gen) / r.numel() ** 0.5 for r in ref]
loss_ref = sum((r * c).sum() for r, c in zip(ref[:7], cots[:7]))
loss_ref.backward()
loss = sum((o * c.cuda()).sum() for o, c in zip(outs[:7], cots[:7]))
t0 = time.time()
loss.backward()
torch.cuda.synchronize()
print("G bwd ok", time.time() - t0, "kernel_status", _lib.kernel_status())
worst = []
for k, p in G.named_parameters():
    worst.append((rel(p.grad, pg[k].grad), k, float(pg[k].grad.norm())))
if os.environ.get('ORDER', '0') == '1':
    for w in worst:
        if w[1].endswith('weight'): print("  grad rel %.2e  %s  (ref norm %.3e)" % w)
worst.sort(reverse=True)
for w in worst[:10]:
    print("  grad rel %.2e  %s  (ref norm %.3e)" % w)
tot_a = torch.cat([p.grad.flatten().cpu() for _, p in G.named_parameters()])
tot_b = torch.cat([pg[k].grad.flatten() for k, _ in G.named_parameters()])
print("  all-params grad rel", rel(tot_a, tot_b))
# D
mp.MASK_HOOK = None
x = cu[0]
dl = D(x)
pd = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
dr = mp.discriminator(pd, b["img"])
print("D logits rel", rel(dl, dr))
cd = torch.randn(dr.shape, generator=gen)
(dr * cd).sum().backward()
(dl * cd.cuda()).sum().backward()
torch.cuda.synchronize()
for k, p in D.named_parameters():
    print("  D grad rel %.2e %s" % (rel(p.grad, pd[k].grad), k))
print("kernel_status", _lib.kernel_status(), "launches", _lib.launch_count())
