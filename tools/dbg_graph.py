import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import step as ostep
from tpgan_b200 import D_and_G_model as M, config, _lib
from tpgan_b200.train_step import TPGANTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
G = M.Generator(64, 347, False, False).cuda(); D = M.Discriminator(False).cuda()
tr = TPGANTrainer(G, D, B, use_dropout=True, use_graphs=True)
b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}
for i in range(5):
    l0 = _lib.launch_count(); torch.cuda.synchronize(); t0 = time.time()
    tr.step(b, read_metrics=False)
    torch.cuda.synchronize()
    print(i, "launch calls", _lib.launch_count() - l0, "wall %.4f" % (time.time() - t0))
