"""Drop-in for the reference's FeatureExtract.py:5-41: FeatureExtractModel(base_model_name, num_of_output_classes,
use_pretrained, **kwargs).forward(x) -> base model output ((logits, FC0 feature) for the ResNet, (locations, classifications)
for MobileNetV2).  The 'resnet' base is the one on the TP-GAN step's path (the identity-preserving loss).  For 'mobilenetv2'
the reference reads `self.base_model.FC[-1].in_features` (FeatureExtract.py:33) of a class that has no `FC` attribute, so
that branch cannot be constructed there; here the base is built and the head the reference describes -
Sequential(Dropout(0.2), Linear(1280, classes)) behind the (constructed but unused, MobileNetV2.py:174) average pool - is
attached as `.base_model.FC`; forward() is MobileNetV2.forward, which never uses it - exactly what the reference's forward
would do.  As in the reference, the final FC is replaced by a fresh Linear(in_features, num_of_output_classes) - the
reference reads `.FC.in_features` of an nn.Sequential there (FeatureExtract.py:31), which cannot work; the intent is kept."""
from __future__ import annotations

import torch.nn as nn

from .ResNet import ResNet18


class FeatureExtractModel(nn.Module):
    def __init__(self, base_model_name="resnet", num_of_output_classes=1000, use_pretrained=False, **kwargs):
        super().__init__()
        self.base_model_name = base_model_name.lower()
        self.num_of_output_classes = num_of_output_classes
        if self.base_model_name == "resnet":
            self.base_model = ResNet18(**kwargs)
            in_features = self.base_model.FC[0].in_features
            self.base_model.FC = nn.Sequential(nn.Linear(in_features, num_of_output_classes))
        elif self.base_model_name == "mobilenetv2":
            from .MobileNetV2 import MobileNetV2
            self.base_model = MobileNetV2(**kwargs)
            self.base_model.FC = nn.Sequential(nn.Dropout(p=0.2), nn.Linear(1280, num_of_output_classes))
        else:
            raise ValueError("FeatureExtractModel supports 'resnet' (ResNet18) or 'mobilenetv2'")

    def forward(self, x):
        return self.base_model(x)
