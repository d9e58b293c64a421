"""Times the per-step housekeeping launches of the generator (gradient export, Adam, weight re-pack) with CUDA events."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import step as ostep
from tpgan_b200 import D_and_G_model as M, config
from tpgan_b200.train_step import TPGANTrainer

B = 4
torch.manual_seed(0)
G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"]).cuda()
D = M.Discriminator(config.D["use_batchnorm"]).cuda()
tr = TPGANTrainer(G, D, B)
b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}
tr.step(b)
flush = torch.zeros(64 * 1024 * 1024, device="cuda")


def t(name, fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:28s} {ts[len(ts)//2]:.3f} ms")


s = tr.g_set
t("bias_tab", s.bias_tab.run)
t("utrans_tab", s.utrans_tab.run)
t("unpack_tab", s.unpack_tab.run)
t("export singles", lambda: [L.export_grad(accumulate=False) for L in s.single])
t("adam G", lambda: tr.flat_g.adam(1e-4))
t("pack_tab", s.pack_tab.run)
t("ptrans_tab", s.ptrans_tab.run)
t("repack singles", lambda: [L.repack() for L in s.single])
print("single layers:", [L.name for L in s.single])
