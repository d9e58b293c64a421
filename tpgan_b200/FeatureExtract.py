"""Drop-in for the reference's FeatureExtract.py:5-41: FeatureExtractModel(base_model_name, num_of_output_classes,
use_pretrained, **kwargs).forward(x) -> base model output ((logits, FC0 feature) for the ResNet).  Only the 'resnet' base
is on the TP-GAN step's path (the identity-preserving loss); 'mobilenetv2' (BASELINE config 5) is out of scope of round 1
and raises.  As in the reference, the final FC is replaced by a fresh Linear(in_features, num_of_output_classes) - the
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
            raise NotImplementedError("the MobileNetV2 base (Pretrain path, BASELINE config 5) is not part of this build")
        else:
            raise ValueError("FeatureExtractModel supports 'resnet' (ResNet18) or 'mobilenetv2'")

    def forward(self, x):
        return self.base_model(x)
