"""Writes tests/golden/*.pt from the LIVE reference modules (oracle/reference.py: /root/reference + the 3-fix shim).
Run in the build container (the reference tree does not exist on the GPU box):  python tools/make_golden.py
Inputs are regenerated from seeds (oracle.step.make_batch) and weights from torch.manual_seed(0) + the reference
constructors, so only outputs are stored."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import reference, step as ostep

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
torch.set_num_threads(8)
G, D = reference.build_models(0)
reference.self_check(G, D)
names = ("img", "left_eye", "right_eye", "nose", "mouth", "z")
gold = {}
# ---- forward, batch 1 (seed 1234) through the reference modules themselves
b = ostep.make_batch(1)
with torch.no_grad():
    outs = G(*[b[k] for k in names], False)
    gold["g_forward_b1"] = [o.clone() for o in outs]
    gold["d_forward_b1"] = D(b["img"]).clone()
    ns = reference.load()
    fuser = ns.DG.LocalFuser()
    gold["fuser_b1"] = fuser(b["left_eye"], b["right_eye"], b["nose"], b["mouth"]).clone()
# ---- the oracle step, batch 2, on the reference modules (losses + gradient fingerprints, no optimiser step)
b2 = ostep.make_batch(2)
Gc = lambda bb: G(*[bb[k] for k in names], False)
g_out = Gc(b2)
ld, md = ostep.d_loss(D, g_out[0].detach(), b2)
gd = torch.autograd.grad(ld, list(D.parameters()))
lg, mg = ostep.g_loss(g_out, D(g_out[0]), b2)
gg = torch.autograd.grad(lg, list(G.parameters()))
gold["step_b2_metrics"] = {k: float(v) for k, v in {**md, **mg}.items()}
gold["step_b2_g_gradnorm"] = {n: float(g.norm()) for (n, _), g in zip(G.named_parameters(), gg)}
gold["step_b2_d_gradnorm"] = {n: float(g.norm()) for (n, _), g in zip(D.named_parameters(), gd)}
keep = ["global_pathway.decoded_img128.0.weight", "global_pathway.conv6.0.bias", "feature_predict.fc.bias",
        "local_pathway_nose.local_img.0.weight", "global_pathway.deconv_128.0.weight"]
gold["step_b2_g_grads"] = {n: g.clone() for (n, _), g in zip(G.named_parameters(), gg) if n in keep}
gold["step_b2_d_grads"] = {n: g.clone() for (n, _), g in zip(D.named_parameters(), gd) if n.endswith("bias") or n == "model.7.0.weight"}
# ---- process() crop boxes from the reference data pipeline for a few landmark sets (incl. out-of-image)
import numpy as np
from PIL import Image
lms = np.array([ostep.MEAN_LANDMARKS, ostep.MEAN_LANDMARKS + 2.6, ostep.MEAN_LANDMARKS - 2.2,
                [[3.2, 5.9], [120.7, 4.1], [64.0, 64.0], [10.5, 125.5], [120.2, 126.9]]], dtype=np.float32)
rng = np.random.RandomState(0)
img = (rng.rand(128, 128, 3) * 255).astype(np.uint8)
crops = []
for lm in lms:
    parts = ns.DataAndDataset.process(Image.fromarray(img), lm.copy())
    crops.append({k: torch.from_numpy(np.asarray(v).copy()) for k, v in parts.items()})
gold["process_landmarks"] = torch.from_numpy(lms)
gold["process_image_u8"] = torch.from_numpy(img)
gold["process_crops"] = crops
torch.save(gold, os.path.join(OUT, "reference_golden.pt"))
print("wrote", os.path.join(OUT, "reference_golden.pt"), os.path.getsize(os.path.join(OUT, "reference_golden.pt")) / 1e6, "MB")
print(gold["step_b2_metrics"])
print(list(crops[0].keys()), [tuple(v.shape) for v in crops[0].values()])
