"""Developer tool: trace the TP-GAN trainer (tf32 or bf16) on the CPU with a stub library (every C-ABI call returns 0,
nothing is computed) to catch host-side tracing bugs without a GPU.  Not part of the product or the tests.
usage: python tools/dry_trace_gan.py [batch] [tf32|bf16] [identity] [bn] [exact]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tpgan_b200 import _lib, ops

CALLS = {}


class _Stub:
    def __getattr__(self, name):
        def f(*a):
            CALLS[name] = CALLS.get(name, 0) + 1
            if name == "tpgan_get_deterministic":
                return 0
            return b"" if name == "tpgan_last_error" else 0
        return f


_lib.load = lambda: _Stub()
ops._stream = lambda: 0
ops._ptr = lambda t: None if t is None else t.data_ptr()
torch.Tensor.pin_memory = lambda self: self
ops.softmax_ce = lambda *a, **k: CALLS.__setitem__("tpgan_softmax_ce", CALLS.get("tpgan_softmax_ce", 0) + 1)
ops.adam_step_dev = lambda *a, **k: CALLS.__setitem__("tpgan_adam_step_dev", CALLS.get("tpgan_adam_step_dev", 0) + 1)

from tpgan_b200 import D_and_G_model as M, config
from tpgan_b200.train_step import TPGANTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dtype = sys.argv[2] if len(sys.argv) > 2 else "tf32"
torch.manual_seed(0)
G = M.Generator(config.G["zdim"], config.G["num_classes"], ("bn" in sys.argv) or config.G["use_batchnorm"], config.G["use_residual_block"])
D = M.Discriminator(config.D["use_batchnorm"])
ident = None
if "identity" in sys.argv:
    from tpgan_b200.FeatureExtract import FeatureExtractModel
    from tpgan_b200.ResNet import BasicBlock
    ident = FeatureExtractModel("resnet", config.G["num_classes"], residualBlock=BasicBlock, feature_layer_dim_before_FC=256).eval()
tr = TPGANTrainer(G, D, B, device="cpu", use_dropout=True, dtype=dtype, identity_net=ident, exact="exact" in sys.argv)
print("plan: fwd", len(tr.plan.fwd), "bwd", len(tr.plan.bwd), "layers", len(tr.plan.layers), "arena MB", tr.arena.total / 2**20)
b = dict(img=torch.rand(B, 3, 128, 128), img_frontal=torch.rand(B, 3, 128, 128), img64_frontal=torch.rand(B, 3, 64, 64),
         img32_frontal=torch.rand(B, 3, 32, 32), landmarks=torch.rand(B, 5, 2) * 100, z=torch.rand(B, 64),
         label=torch.zeros(B, dtype=torch.int64), gp_alpha=torch.rand(B))
CALLS.clear()
tr.step(b, read_metrics=False)
print("one step:", sum(CALLS.values()), "C-ABI calls")
for k, v in sorted(CALLS.items(), key=lambda kv: -kv[1]):
    print(f"  {k}: {v}")
if dtype == "bf16":
    print("casts fwd:", [f.label for f in tr.plan.fwd if getattr(f, "kind", "") == "cast16"])
    print("casts bwd:", [f.label for f in tr.plan.bwd if getattr(f, "kind", "") == "cast16"])
