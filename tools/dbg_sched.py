"""Times every entry of the step schedule with CUDA events (eager) and aggregates by kind."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import step as ostep
from tpgan_b200 import D_and_G_model as M, _lib
from tpgan_b200.train_step import TPGANTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
G = M.Generator(64, 347, False, False).cuda(); D = M.Discriminator(False).cuda()
tr = TPGANTrainer(G, D, B, use_dropout=True)
b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}
for _ in range(2):
    tr.step(b, read_metrics=False)
sch = tr._schedule(True)
torch.cuda.synchronize()
ev = []
for f in sch:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); f(); e.record()
    ev.append((f, s, e))
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for f, s, e in ev:
    t = s.elapsed_time(e)
    tot += t
    kind = getattr(f, "kind", None)
    if kind is None:
        kind = getattr(f, "__qualname__", type(f).__name__)
        code = getattr(f, "__code__", None)
        if code is not None and "<lambda>" in kind:
            kind = f"lambda@{os.path.basename(code.co_filename)}:{code.co_firstlineno}"
    agg[kind][0] += t; agg[kind][1] += 1
print("entries", len(sch), "sum of entry times %.2f ms" % tot)
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%8.3f ms  n=%4d  %s" % (t, n, k))
