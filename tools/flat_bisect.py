"""Developer tool (GPU box): run the generator's forward launch list at batch B twice per conv launch - multi-tap GEMM kernel
vs flat-slab kernel (TPGAN_FLATCONV is read per call) - and report every launch whose results differ beyond tf32 noise."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TPGAN_FLATCONV"] = "0"
from tpgan_b200 import _lib, ops  # noqa: E402

CHUNKS = []
_orig = ops.Arena.alloc


def _alloc(self, shape, dtype=torch.float32):
    before = self.chunk
    t = _orig(self, shape, dtype)
    if self.chunk is not before:
        CHUNKS.append(self.chunk)
    return t


ops.Arena.alloc = _alloc
from oracle import step as ostep  # noqa: E402
from tpgan_b200 import D_and_G_model as M, config  # noqa: E402
from tpgan_b200.train_step import TPGANTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"]).cuda()
D = M.Discriminator(config.D["use_batchnorm"]).cuda()
tr = TPGANTrainer(G, D, B, use_graphs=False)
b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}
tr.step(b, optimize=False)
torch.cuda.synchronize()
print("chunks", len(CHUNKS), "MB", sum(c.numel() for c in CHUNKS) >> 20)
tr.load_inputs(b)
tr._stage()
for i, f in enumerate(tr.plan.fwd):
    os.environ["TPGAN_FLATCONV"] = "0"
    if getattr(f, "kind", None) not in ("tapgemm", "rowconv"):
        f()
        continue
    f()
    torch.cuda.synchronize()
    snap = [c.clone() for c in CHUNKS]
    os.environ["TPGAN_FLATCONV"] = "1"
    f()
    torch.cuda.synchronize()
    if _lib.last_conv_kernel() != "flatconv":
        continue
    worst, nbad, ref = 0.0, 0, 0.0
    for c, s in zip(CHUNKS, snap):
        ci, si = c.view(torch.int32), s.view(torch.int32)
        ne = ci != si
        if bool(ne.any()):
            d = (c.view(torch.float32)[ne] - s.view(torch.float32)[ne]).abs()
            worst = max(worst, float(d.max()))
            nbad += int((d > 1e-3).sum())
            ref = max(ref, float(s.view(torch.float32)[ne].abs().max()))
    print(f"{i:4d} {getattr(f, 'label', '?')[:70]:70s} maxdiff {worst:.3e} (ref max {ref:.3e}) elements>1e-3: {nbad}", flush=True)
    for c, s in zip(CHUNKS, snap):
        c.copy_(s)
    del snap
assert _lib.kernel_status() == 0
