"""Checkpoint compatibility of the fused G/D training step: the reference's helpers save_model /
save_optimizer (UtilityMethods.py:58-103) applied to the trainer's modules and fused Adam state produce files that the
oracle modules and stock torch.optim.Adam load, and that a fresh trainer resumes from."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gd_trainer_checkpoint_round_trip(tmp_path):
    from oracle import step as ostep
    from tpgan_b200 import D_and_G_model as M, config
    from tpgan_b200.train_step import TPGANTrainer
    from tpgan_b200.UtilityMethods import save_model, save_optimizer
    torch.manual_seed(0)
    G = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"]).cuda()
    D = M.Discriminator(config.D["use_batchnorm"]).cuda()
    tr = TPGANTrainer(G, D, 1)
    b = {k: v.cuda() for k, v in ostep.make_batch(1).items()}
    tr.step(b)
    tr.step(b)
    gp, dp = save_model(G, str(tmp_path / "g"), 7), save_model(D, str(tmp_path / "d"), 7)
    op = save_optimizer(tr.optimizer_d, D, str(tmp_path / "d"), 7)
    sd_g, sd_d = torch.load(gp, map_location="cpu"), torch.load(dp, map_location="cpu")
    assert len(sd_g) == 328 and len(sd_d) == 20                                    # the reference's key counts
    ck = torch.load(op, map_location="cpu")
    assert set(ck) == {"optimizer", "model", "epoch"} and ck["epoch"] == 7
    params = [torch.nn.Parameter(v.clone()) for v in sd_d.values()]
    adam = torch.optim.Adam(params, lr=1.0)
    adam.load_state_dict(ck["optimizer"])                                          # stock Adam accepts the fused state
    st = adam.state[params[0]]
    assert float(st["step"]) == 2.0 and float(st["exp_avg_sq"].sum()) > 0 and adam.param_groups[0]["lr"] == ostep.LEARNING_RATE
    # resume: fresh modules + trainer, state loaded from the files, next step identical
    torch.manual_seed(5)
    G2 = M.Generator(config.G["zdim"], config.G["num_classes"], config.G["use_batchnorm"], config.G["use_residual_block"]).cuda()
    D2 = M.Discriminator(config.D["use_batchnorm"]).cuda()
    G2.load_state_dict(torch.load(gp))
    D2.load_state_dict(torch.load(dp))
    tr2 = TPGANTrainer(G2, D2, 1)
    tr2.optimizer_d.load_state_dict(torch.load(op)["optimizer"])
    tr2.optimizer_g.load_state_dict(tr.optimizer_g.state_dict())
    assert torch.equal(tr2.flat_d.m, tr.flat_d.m) and torch.equal(tr2.flat_d.v, tr.flat_d.v)
    assert int(tr2.flat_d.step_dev) == int(tr.flat_d.step_dev) == 2
    m1, m2 = tr.step(b), tr2.step(b)
    for k in m1:
        assert abs(m1[k] - m2[k]) <= 1e-4 * abs(m1[k]) + 1e-6, (k, m1[k], m2[k])


def test_gd_training_loop_runs_and_checkpoints(tmp_path):
    """tpgan_b200.train.main: the reference-style G/D loop (uint8 inputs normalised / cropped / pyramided on the device,
    prefetch of the next batch, save_model / save_optimizer per epoch) on synthetic batches."""
    import math

    from tpgan_b200 import train as T
    hist = T.main(["--epochs", "1", "--steps-per-epoch", "3", "--batch", "1", "--log-every", "2", "--save-dir", str(tmp_path)])
    assert len(hist) == 3 and all(math.isfinite(v) for m in hist for v in m.values())
    for net in ("G", "D"):
        assert (tmp_path / net / "model_epoch_0.pth").exists() and (tmp_path / net / "optimizer_epoch_0.pth").exists()
    assert len(torch.load(tmp_path / "G" / "model_epoch_0.pth", map_location="cpu")) == 328
