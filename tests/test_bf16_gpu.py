"""GPU parity of the bf16 operand mode of the fused G+D training step (BASELINE.json configs[2]: "bf16"; north-star
tolerance for bf16: rel 1e-2, stated per quantity below).

What "bf16" means here: activations, activation gradients and weights are rounded to bf16 where a convolution
reads them (tcgen05.mma kind::f16), accumulation is fp32 in TMEM, master weights / optimizer / residual adds / gradient
accumulation / masks / losses stay fp32.  Two comparisons:
  * against the oracle with the SAME operand rounding (oracle.model_port.EMULATE_BF16) under identical activation masks:
    the arithmetic then differs by fp32 summation order only, so this check is sharp - it is the one that would catch a
    stale bf16 copy, a missing cast or a mis-routed operand, all of which hide inside 1e-2;
  * against the plain fp32 oracle: the bf16 format's own deviation, bounded at the north-star's 1e-2 for the loss scalars
    and measured / bounded for forward tensors and gradients."""
import json
import os

import pytest
import torch

from test_model_gpu import GOLD, NAMES, _grad_errors, _models, _oracle_step_grads, rel

pytestmark = pytest.mark.gpu


def _record(name, data):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "bf16_parity.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **data}) + "\n")


def test_step_bf16_vs_oracle():
    from oracle import model_port as mp, step as ostep
    from tpgan_b200 import _lib
    from tpgan_b200.train_step import TPGANTrainer
    B = 2
    G, D, sg, sd = _models(False)
    b = ostep.make_batch(B)
    tr = TPGANTrainer(G, D, B, dtype="bf16")
    m = tr.step({k: v.cuda() for k, v in b.items()}, optimize=False)
    torch.cuda.synchronize()
    assert _lib.kernel_status() == 0
    assert (tr.boxes.cpu().numpy() == ostep.crop_boxes(b["landmarks"].numpy())).all()
    rec = {"metrics": m}
    # ---- forward: the generator's image against the golden record of the live reference (fp32)
    gold = torch.load(GOLD, weights_only=False)
    # ---- (1) same-rounding oracle, identical masks: sharp
    mp.EMULATE_BF16 = True
    try:
        ref_e, gg, gd = _oracle_step_grads(b, sg, sd, tr)
    finally:
        mp.EMULATE_BF16 = False
    rec["metrics_emulated"] = ref_e
    eg, each_g = _grad_errors(G, gg)
    ed, each_d = _grad_errors(D, gd)
    rec.update(emul_g_overall=eg, emul_g_worst=max(each_g.items(), key=lambda kv: kv[1]), emul_d_overall=ed,
               emul_d_worst=max(each_d.items(), key=lambda kv: kv[1]))
    # ---- (2) plain fp32 oracle, identical masks, and un-masked
    ref, gg, gd = _oracle_step_grads(b, sg, sd, tr)
    fg, feach_g = _grad_errors(G, gg)
    fd, feach_d = _grad_errors(D, gd)
    rec.update(metrics_fp32=ref, fp32_g_overall=fg, fp32_g_worst=max(feach_g.items(), key=lambda kv: kv[1]), fp32_d_overall=fd,
               fp32_d_worst=max(feach_d.items(), key=lambda kv: kv[1]))
    _, gg, gd = _oracle_step_grads(b, sg, sd, None)
    ug, _ = _grad_errors(G, gg)
    ud, _ = _grad_errors(D, gd)
    rec.update(unmasked_g_overall=ug, unmasked_d_overall=ud)
    _record("step_b2", rec)
    for k, v in ref_e.items():     # same rounding points: the loss scalars agree far below the format's precision
        assert abs(m[k] - v) <= 2e-3 * abs(v) + 1e-4, ("emulated", k, m[k], v)
    for k, v in ref.items():       # against fp32: the north-star's bf16 tolerance
        assert abs(m[k] - v) <= 1e-2 * abs(v) + 1e-3, ("fp32", k, m[k], v)
    assert eg < 1e-2 and ed < 3e-2, (eg, ed)                  # G: sharp; D: the penalty's tangent pass rounds elsewhere
    assert fg < 3e-2 and fd < 5e-2, (fg, fd)
    assert ug < 0.3 and ud < 0.3, (ug, ud)


def test_forward_bf16_tensors():
    """Stored activations of the bf16 step against the fp32 oracle: the generator image and the critic logits."""
    from oracle import model_port as mp, step as ostep
    from tpgan_b200.train_step import TPGANTrainer
    B = 2
    G, D, sg, sd = _models(False)
    b = ostep.make_batch(B)
    tr = TPGANTrainer(G, D, B, dtype="bf16")
    tr.step({k: v.cuda() for k, v in b.items()}, optimize=False)
    torch.cuda.synchronize()
    fake = tr.fake.act.to_nchw().cpu()
    outs = {}
    for emu in (False, True):
        mp.EMULATE_BF16 = emu
        try:
            with torch.no_grad():
                o = mp.generator(sg, *[b[k] for k in NAMES])
                dl = mp.discriminator(sd, o[0])
        finally:
            mp.EMULATE_BF16 = False
        outs[emu] = (o, dl)
    e_fp32, e_emu = rel(fake, outs[False][0][0]), rel(fake, outs[True][0][0])
    logits = tr.logits.act.to_nchw().reshape(B, -1).cpu()
    l_fp32, l_emu = rel(logits, outs[False][0][1]), rel(logits, outs[True][0][1])
    twin = tr.fake.act.twin().float().permute(0, 3, 1, 2).cpu()
    _record("forward_b2", dict(fake_vs_fp32=e_fp32, fake_vs_emulated=e_emu, logits_vs_fp32=l_fp32, logits_vs_emulated=l_emu,
                               twin_vs_fp32copy=rel(twin, fake)))
    assert e_emu < 2e-3 and l_emu < 2e-3, (e_emu, l_emu)       # same rounding points: sign-flip / summation-order level
    assert e_fp32 < 2e-2 and l_fp32 < 2e-2, (e_fp32, l_fp32)   # the format: ~60 stacked convolutions x 2^-9
    assert rel(twin, fake) < 4e-3                              # the bf16 copy is the rounded fp32 copy


def test_bf16_training_trajectory_and_graphs():
    """Four optimizer steps in bf16: CUDA-graph replay equals eager launches bit for bit in deterministic mode (the twins,
    casts and bf16 re-packs are all inside the captured schedule), the losses stay finite and move."""
    from oracle import step as ostep
    from tpgan_b200 import _lib
    from tpgan_b200.train_step import TPGANTrainer
    B = 2
    b = {k: v.cuda() for k, v in ostep.make_batch(B).items()}

    def run(graphs):
        G, D, _, _ = _models(False)
        tr = TPGANTrainer(G, D, B, use_graphs=graphs, dtype="bf16")
        ms = [tr.step(b, optimize=True) for _ in range(4)]
        torch.cuda.synchronize()
        return ms, tr.flat_g.data.clone(), tr.flat_d.data.clone()

    prev = _lib.set_deterministic(True)
    try:
        me, pg_e, pd_e = run(False)
        mg, pg_g, pd_g = run(True)
    finally:
        _lib.set_deterministic(prev)
    assert _lib.kernel_status() == 0
    assert torch.equal(pg_e, pg_g) and torch.equal(pd_e, pd_g), (rel(pg_g, pg_e), rel(pd_g, pd_e))
    assert all(v == v and abs(v) < 1e4 for m in me for v in m.values()), me
    assert me[-1]["g_total"] < me[0]["g_total"]        # the same batch four times: the generator loss goes down
    _record("trajectory", dict(first=me[0], last=me[-1]))
