"""Developer tool: trace the MobileNetV2 plan / PretrainTrainer schedule on the CPU with a stub library (every C-ABI call
returns 0, nothing is computed) to catch host-side tracing bugs without a GPU.  Not part of the product or the tests."""
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
            return b"" if name == "tpgan_last_error" else 0
        return f


_lib.load = lambda: _Stub()
ops._stream = lambda: 0
ops._ptr = lambda t: None if t is None else t.data_ptr()

from tpgan_b200.MobileNetV2 import MobileNetV2
from tpgan_b200.pretrain_step import PretrainTrainer

torch.manual_seed(0)
m = MobileNetV2()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tr = PretrainTrainer(m, B, device="cpu")
print("plan: fwd launches", len(tr.plan.fwd), "bwd", len(tr.plan.bwd), "conv layers", len(tr.plan.layers), "aux", len(tr.plan.aux),
      "n points", tr.n, "act MB", tr.plan.bytes / 2**20)
CALLS.clear()
x = torch.rand(B, 3, 128, 128)
tr.step(x, torch.rand(B, 8) * 128, None, read_metrics=False)
print("one step:", sum(CALLS.values()), "C-ABI calls")
for k, v in sorted(CALLS.items(), key=lambda kv: -kv[1]):
    print(f"  {k}: {v}")
# every parameter must have a gradient writer
named = dict(m.named_parameters())
print("params", len(named), "flat floats", tr.flat.total)
