"""Writes tests/golden/pretrain_golden.pt from the LIVE reference (/root/reference/MobileNetV2.py, unmodified):
  * MobileNetV2 built under torch.manual_seed(0): parameter fingerprints, train-mode forward on a seeded batch of 2
    (outputs + running statistics afterwards), eval-mode forward;
  * MultiTaskLoss on seeded random inputs (n = 394), with torch.multinomial replaced by the explicit-key rule of
    oracle/pretrain_port.py (documented there): loss values, assignments, gradients;
  * the Temp.py known answer.
Run in the build container (needs /root/reference):  python tools/make_golden_pretrain.py"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("TPGAN_REFERENCE_DIR", "/root/reference"))
import MobileNetV2 as R  # noqa: E402

from oracle.pretrain_port import make_batch  # noqa: E402


def loss_case(seed, n=394):
    g = torch.Generator().manual_seed(seed)
    loc = (torch.rand((1, n, 2), generator=g) * 140 - 6).clamp_min(0)
    cls = torch.randn((1, n, 5), generator=g)
    true = torch.tensor([[39.48, 40.28, 85.96, 38.7, 63.64, 63.65, 64.78, 89.32]]) + torch.rand((1, 8), generator=g) * 6 - 3
    u = torch.rand((1, n), generator=g)
    return loc, cls, true, u


def main():
    out = {}
    torch.manual_seed(0)
    net = R.MobileNetV2()
    out["param_sums"] = {k: float(v.double().sum()) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    x, _, _ = make_batch(2, seed=11)
    net.train()
    with torch.no_grad():
        loc, cls = net(x)
    out["train_loc"], out["train_cls"] = loc.clone(), cls.clone()
    out["running"] = {k: v.clone() for k, v in net.state_dict().items() if "running_" in k and k.startswith(("conv1", "conv2"))}
    net.eval()
    with torch.no_grad():
        loc, cls = net(x)
    out["eval_loc"], out["eval_cls"] = loc.clone(), cls.clone()
    cases = []
    for seed in range(4):
        loc, cls, true, u = loss_case(seed)
        loc.requires_grad_(True), cls.requires_grad_(True)
        orig = torch.multinomial

        def keyed(w, m, replacement=False, u=u):
            key = torch.where(w > 0, u[0], torch.full_like(u[0], float("inf")))
            return torch.sort(key, stable=True)[1][:m]
        torch.multinomial = keyed
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                L = R.MultiTaskLoss()
                val = L(loc, cls, true, (128, 128))
                _, labels = L.get_positive_samples_and_classification_tensor(loc, true)
        finally:
            torch.multinomial = orig
        gl, gc = torch.autograd.grad(val, [loc, cls])
        cases.append(dict(seed=seed, loss=float(val), labels=labels.clone(), dloc=gl.clone(), dcls=gc.clone()))
    out["loss_cases"] = cases
    out["temp_py_total_loss"] = 0.8939134478569031      # printed by /root/reference/Temp.py
    path = os.path.join(ROOT, "tests", "golden", "pretrain_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
