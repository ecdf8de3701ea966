"""ResNet-50 trunks that return the four stage feature maps (producer of the token builder).

Mirrors models/resnet50ssl.py of the reference (ResNetTrunk :12-27, ResNetTrunkByScale :30-45,
resnet50FeatureExtractor :60-79).  These classes only hold the parameters (state_dict schema); the forward of the
bf16 path runs the BN-folded trunk on the package's own convolution kernel (trunk_convs.py, csrc/conv_tcgen05.cu).  Weight download (Lunit TCGA SSL
checkpoints, :48-57) needs network access and is only attempted when `pretrained=True`; a
local file of the reference's name is used when present.
"""
from __future__ import annotations

import os

import torch
from torchvision.models.resnet import Bottleneck, ResNet


class ResNetTrunk(ResNet):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.fc

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class ResNetTrunkByScale(ResNet):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.fc

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x0 = self.layer1(x)
        x1 = self.layer2(x0)
        x2 = self.layer3(x1)
        x3 = self.layer4(x2)
        return [x0, x1, x2, x3]


def get_pretrained_url(key):
    prefix = "https://github.com/lunit-io/benchmark-ssl-pathology/releases/download/pretrained-weights"
    registry = {"BT": "bt_rn50_ep200.torch", "MoCoV2": "mocov2_rn50_ep200.torch", "SwAV": "swav_rn50_ep200.torch"}
    return f"{prefix}/{registry.get(key)}", registry.get(key)


def _load_pretrained(model, progress, key):
    url, filename = get_pretrained_url(key)
    if os.path.exists(filename):
        state_dict = torch.load(filename, map_location="cpu")
    else:
        state_dict = torch.hub.load_state_dict_from_url(url, progress=progress)
        torch.save(state_dict, filename)
    model.load_state_dict(state_dict)
    return model


def resnet50FeatureExtractor(pretrained, progress, key, **kwargs):
    model = ResNetTrunkByScale(Bottleneck, [3, 4, 6, 3], **kwargs)
    return _load_pretrained(model, progress, key) if pretrained else model


def resnet50(pretrained, progress, key, **kwargs):
    model = ResNetTrunk(Bottleneck, [3, 4, 6, 3], **kwargs)
    return _load_pretrained(model, progress, key) if pretrained else model
