"""ORACLE (test infrastructure, never shipped on the product path).

Loads the reference TP-GAN modules from /root/reference *in process* and applies the minimal fixes without which the
reference cannot construct its own Generator/Discriminator:

  F1  ModificationLayer.py:103,191 hand the nn.Module (not .weight) to weight_initialization (:26-52) and use the
      removed-in-spirit aliases kaiming_normal/xavier_normal  -> unwrap to .weight, call the in-place initialisers.
  F2  ModificationLayer.py:154 appends activation=None into nn.Sequential                     -> drop None entries.
  F3  D_and_G_model.py:268 dim128 omits the 3 channels of I128 that forward() concatenates (:323)  -> + 3.
  (F4 ModificationLayer.py:146 calls activation() - only reached with use_batchnorm=True, which config.py:63,68 turn
      off; patched the same way so the BN path is at least constructible.)

Nothing from the reference is copied into this repository: the modules are imported from their read-only location,
two module attributes are rebound and ONE source line is patched in memory before exec.  Only tests/, tools that write
golden fixtures, __graft_entry__.smoke() and bench.py's cpu baseline may import this file.

The reference tree is absent on the GPU box; `available()` says whether it can be loaded.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_DIR = os.environ.get("TPGAN_REFERENCE_DIR", "/root/reference")
_F3_OLD = "dim128 = self.conv0.out_channels + self.deconv_128.out_channels"
_F3_NEW = "dim128 = self.conv0.out_channels + self.deconv_128.out_channels + 3"

_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "D_and_G_model.py"))


def _weight_initialization(weight, init, activation):
    """Restates ModificationLayer.py:26-52 with fix F1."""
    if init is None:
        return
    w = weight.weight if isinstance(weight, nn.Module) else weight
    if init == "kaiming":
        a = activation.negative_slope if hasattr(activation, "negative_slope") else 0
        nn.init.kaiming_normal_(w, a=a)
    elif init == "xavier":
        nn.init.xavier_normal_(w)


def _batchnorm_and_activation_layer(specific_channels, activation, use_batchnorm):
    """Restates ModificationLayer.py:125-156 with fixes F2 and F4."""
    layers = []
    if use_batchnorm:
        if isinstance(activation, (nn.Sigmoid, nn.Tanh)):
            layers += [activation, nn.BatchNorm2d(specific_channels)]
        else:
            layers += [nn.BatchNorm2d(specific_channels), activation]
    else:
        layers.append(activation)
    return [l for l in layers if l is not None]


def load():
    """Returns a namespace with the shimmed reference modules: .ML (ModificationLayer), .DG (D_and_G_model),
    .config, .DataAndDataset (process()), .UtilityMethods."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    sys.dont_write_bytecode = True
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import ModificationLayer as ML  # noqa: N811
    ML.weight_initialization = _weight_initialization
    ML._batchnorm_and_activation_layer = _batchnorm_and_activation_layer
    import config as ref_config
    import UtilityMethods
    src = open(os.path.join(REFERENCE_DIR, "D_and_G_model.py"), encoding="utf-8").read()
    assert src.count(_F3_OLD) == 1, "reference changed: F3 patch site not found"
    src = src.replace(_F3_OLD, _F3_NEW)
    DG = types.ModuleType("D_and_G_model_shimmed")
    DG.__file__ = os.path.join(REFERENCE_DIR, "D_and_G_model.py")
    exec(compile(src, DG.__file__, "exec"), DG.__dict__)
    import DataAndDataset
    ns = types.SimpleNamespace(ML=ML, DG=DG, config=ref_config, DataAndDataset=DataAndDataset,
                               UtilityMethods=UtilityMethods)
    _cache["ns"] = ns
    return ns


def build_models(seed: int = 0):
    """Reference G and D with config.py hyper-parameters (config.py:59-68), seeded."""
    ns = load()
    cfg = ns.config
    torch.manual_seed(seed)
    G = ns.DG.Generator(cfg.G["zdim"], cfg.G["num_classes"], cfg.G["use_batchnorm"], cfg.G["use_residual_block"])
    D = ns.DG.Discriminator(cfg.D["use_batchnorm"])
    return G, D


EXPECTED = dict(g_params=137_764_238, d_params=13_354_625, local_params=12_466_627, global_params=87_808_551,
                feature_predict_params=89_179, g_tensors=328, d_tensors=20)


def self_check(G, D) -> None:
    """Invariants measured in the survey."""
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(G) == EXPECTED["g_params"], n(G)
    assert n(D) == EXPECTED["d_params"], n(D)
    assert n(G.local_pathway_nose) == EXPECTED["local_params"]
    assert n(G.global_pathway) == EXPECTED["global_params"]
    assert n(G.feature_predict) == EXPECTED["feature_predict_params"]
    assert len(G.state_dict()) == EXPECTED["g_tensors"] and len(D.state_dict()) == EXPECTED["d_tensors"]
