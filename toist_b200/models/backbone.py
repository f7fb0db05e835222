"""ResNet-50/101 trunk with FrozenBatchNorm (reference models/backbone.py:21-91,165-198 + torchvision resnet).

The modules below are *parameter containers* whose names reproduce the reference's state-dict layout
(`backbone.0.body.layer3.5.conv2.weight`, `...bn2.running_var`, `...downsample.0.weight`); the arithmetic runs in
runtime.BACKBONE on NHWC bf16 implicit-GEMM kernels with the BatchNorm affine folded into the epilogue.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import List, Tuple

import torch
from torch import nn

from .. import kernels as K
from ..runtime import RESNET_BLOCKS
from ..util.misc import NestedTensor
from .position_encoding import build_position_encoding


class FrozenBatchNorm2d(nn.Module):
    """Fixed statistics and affine parameters, kept as buffers (eps = 1e-5 inside the rsqrt)."""

    def __init__(self, n: int):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class ConvWeight(nn.Module):
    """Bias-free convolution weight [Cout, Cin, k, k] (torchvision init: kaiming normal, fan_out, relu)."""

    def __init__(self, cin: int, cout: int, k: int, stride: int = 1):
        super().__init__()
        self.stride = stride
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))

    def reset_parameters(self) -> None:
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int, downsample: bool):
        super().__init__()
        self.conv1 = ConvWeight(inplanes, planes, 1)
        self.bn1 = FrozenBatchNorm2d(planes)
        self.conv2 = ConvWeight(planes, planes, 3, stride)
        self.bn2 = FrozenBatchNorm2d(planes)
        self.conv3 = ConvWeight(planes, planes * 4, 1)
        self.bn3 = FrozenBatchNorm2d(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(ConvWeight(inplanes, planes * 4, 1, stride), FrozenBatchNorm2d(planes * 4))
        self.stride = stride


class ResNetBody(nn.Module):
    """conv1 / bn1 / layer1..layer4 of torchvision's resnet50 / resnet101 (no avgpool / fc: IntermediateLayerGetter
    drops them in the reference, models/backbone.py:71)."""

    def __init__(self, name: str):
        super().__init__()
        if name not in RESNET_BLOCKS:
            raise ValueError(f"backbone {name!r} is outside the TOIST hot path (resnet50 / resnet101)")
        self.arch = name
        self.blocks = RESNET_BLOCKS[name]
        self.conv1 = ConvWeight(3, 64, 7, 2)
        self.bn1 = FrozenBatchNorm2d(64)
        inplanes = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), self.blocks), start=1):
            stride = 1 if li == 1 else 2
            layer = [Bottleneck(inplanes, planes, stride, True)]
            inplanes = planes * 4
            layer += [Bottleneck(inplanes, planes, 1, False) for _ in range(1, n)]
            setattr(self, f"layer{li}", nn.Sequential(*layer))

    def reset_parameters(self) -> None:
        """Same random stream as the reference's `torchvision.models.resnetNN(norm_layer=FrozenBatchNorm2d)`
        (models/backbone.py:87): torchvision draws the default Conv2d / fc initialisations first and then re-draws every
        convolution with kaiming-normal, so the simplest way to consume the generator identically is to let
        torchvision construct a throw-away copy and take its values.  Falls back to plain kaiming-normal."""
        try:
            import torchvision

            tv = getattr(torchvision.models, self.arch)(weights=None, norm_layer=FrozenBatchNorm2d)
            sd = {k: v for k, v in tv.state_dict().items() if not k.startswith("fc.")}
            self.load_state_dict(sd, strict=True)
        except ImportError:  # pragma: no cover
            for m in self.modules():
                if isinstance(m, ConvWeight):
                    m.reset_parameters()

    def conv_bn_pairs(self) -> List[Tuple[str, str]]:
        pairs = [("conv1", "bn1")]
        for li, n in enumerate(self.blocks, start=1):
            for bi in range(n):
                p = f"layer{li}.{bi}."
                pairs += [(p + "conv1", p + "bn1"), (p + "conv2", p + "bn2"), (p + "conv3", p + "bn3")]
                if bi == 0:
                    pairs.append((p + "downsample.0", p + "downsample.1"))
        return pairs


class Backbone(nn.Module):
    """ResNet trunk; stem and layer1 are always frozen, layer2-4 train when `train_backbone` (backbone.py:64-66)."""

    def __init__(self, name: str, train_backbone: bool, return_interm_layers: bool, dilation: bool):
        super().__init__()
        if dilation:
            raise NotImplementedError("dilation (DC5) is outside the TOIST hot path (main.py:99-103 default False)")
        self.body = ResNetBody(name)
        self.body.reset_parameters()
        for pname, p in self.body.named_parameters():
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                p.requires_grad_(False)
        self.train_backbone = train_backbone
        self.return_interm_layers = return_interm_layers
        self.num_channels = 2048
        path = os.environ.get("TOIST_BACKBONE_WEIGHTS")
        if path:  # torchvision resnet state dict (the reference downloads it with pretrained=True)
            sd = torch.load(path, map_location="cpu")
            sd = {k: v for k, v in sd.items() if not k.startswith("fc.")}
            self.body.load_state_dict(sd, strict=True)

    def forward(self, tensor_list: NestedTensor):
        raise NotImplementedError(
            "the trunk has no standalone forward in toist_b200: it runs inside MDETR.forward (runtime.BACKBONE), which "
            "keeps activations NHWC bf16 end to end; call the model, not its backbone")


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)

    def forward(self, tensor_list):
        xs = self[0](tensor_list)
        out, pos = [], []
        for _, x in xs.items():
            out.append(x)
            pos.append(self[1](x).to(x.tensors.dtype))
        return out, pos


def build_backbone(args):
    position_embedding = build_position_encoding(args)
    train_backbone = args.lr_backbone > 0
    return_interm_layers = args.masks
    backbone = Backbone(args.backbone, train_backbone, return_interm_layers, args.dilation)
    model = Joiner(backbone, position_embedding)
    model.num_channels = backbone.num_channels
    return model
