"""Writes tests/golden/step_bn_golden.pt: ONE oracle training step (oracle/step.py: D phase with gradient penalty, G phase,
fixed weights) of the LIVE reference Generator built with use_batchnorm=True (the constructor's default,
D_and_G_model.py:351) against the config's BatchNorm-free Discriminator (config.py:68), on the seeded batch of 2 that the
other BatchNorm goldens use.  Recorded: the 11 loss scalars, per-parameter gradient norm / sum of both networks, the full
BatchNorm affine gradients, D's full gradients (20 small tensors would be 11 M floats - norms only) and the running
statistics of four BatchNorm layers after the step's single generator forward.
Run in the build container (needs /root/reference):  python tools/make_golden_bn_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference as R  # noqa: E402
from oracle import step as ostep  # noqa: E402

RUNNING = ("global_pathway.conv0.0.1.running_mean", "global_pathway.conv0.0.1.running_var",
           "global_pathway.deconv_8.1.running_var", "local_pathway_nose.after_select2.0.1.running_mean")


def main():
    ns = R.load()
    torch.manual_seed(0)
    G = ns.DG.Generator(64, 347, True, False)
    D = ns.DG.Discriminator(False)
    G.train()
    D.train()
    b = ostep.make_batch(2, seed=3)
    opt_g = torch.optim.Adam(G.parameters(), lr=1e-4)
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4)

    def Gc(bb):
        return G(bb["img"], bb["left_eye"], bb["right_eye"], bb["nose"], bb["mouth"], bb["z"], False)

    m = ostep.train_step(Gc, D, list(G.parameters()), list(D.parameters()), opt_g, opt_d, b, step_optim=False)
    # train_step leaves D's gradients of the D phase (zero_grad(set_to_none) only touches the optimizer of the phase) and
    # G's gradients of the G phase in .grad
    sd = G.state_dict()
    gold = dict(metrics=m,
                g_grad_stats={k: (float(p.grad.norm()), float(p.grad.double().sum())) for k, p in G.named_parameters()},
                d_grad_stats={k: (float(p.grad.norm()), float(p.grad.double().sum())) for k, p in D.named_parameters()},
                bn_grads={k: p.grad.clone() for k, p in G.named_parameters() if p.dim() == 1 and ".1." in k and p.numel() <= 512},
                running={k: sd[k].clone() for k in RUNNING},
                num_batches_tracked=int(sd["global_pathway.conv0.0.1.num_batches_tracked"]))
    path = os.path.join(ROOT, "tests", "golden", "step_bn_golden.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes;", m)


if __name__ == "__main__":
    main()
