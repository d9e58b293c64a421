"""Drop-in for the hot-path helpers of the reference's UtilityMethods.py (:14-121): optimizer factory,
set_requires_grad, checkpoint writers (same file names and dict layout, so checkpoints are interchangeable) and the
channel-width helper used by the model constructors.  PIL-based image helpers (:123-164) are data-pipeline code and out
of scope."""
from __future__ import annotations

import os

import torch
from torch import optim

from .config import optimizer_param


def getOptimizer(model_parameters, optimizer_name="SGD"):
    """Returns a zero-argument factory for the named optimizer built from config.optimizer_param, like the reference
    (UtilityMethods.py:14-41; unknown names fall back to SGD)."""
    model_parameters = list(model_parameters)
    lr, wd, mom = optimizer_param["learning_rate"], optimizer_param["weight_decay"], optimizer_param["momentum"]
    table = {
        "SGD": lambda: optim.SGD(model_parameters, lr=lr, weight_decay=wd, momentum=mom,
                                 nesterov=optimizer_param.get("nesterov", False)),
        "Adam": lambda: optim.Adam(model_parameters, lr=lr, weight_decay=wd),
        "RMSprop": lambda: optim.RMSprop(model_parameters, lr=lr, weight_decay=wd, momentum=mom),
        "Adagrad": lambda: optim.Adagrad(model_parameters, lr=lr, weight_decay=wd),
        "Adadelta": lambda: optim.Adadelta(model_parameters, lr=lr, weight_decay=wd),
    }
    return table.get(optimizer_name, table["SGD"])


def set_requires_grad(parameters, isGrad):
    """UtilityMethods.py:43-56."""
    for param in parameters:
        param.requires_grad = isGrad


def save_model(model, dir, epoch):
    """{dir}/model_epoch_{epoch}.pth = state_dict (UtilityMethods.py:58-76)."""
    path = os.path.join(dir, f"model_epoch_{epoch}.pth")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(model.state_dict(), path)
    return path


def save_optimizer(optimizer, model, dir, epoch):
    """{dir}/optimizer_epoch_{epoch}.pth = {'optimizer', 'model', 'epoch'} (UtilityMethods.py:78-103)."""
    path = os.path.join(dir, f"optimizer_epoch_{epoch}.pth")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({"optimizer": optimizer.state_dict(), "model": model.state_dict(), "epoch": epoch}, path)
    return path


def elementwise_multiply_and_cast_to_int(list_x, scalar):
    """UtilityMethods.py:109-121."""
    return [int(v * scalar) for v in list_x]
