"""Parameter containers for the GELAN blocks.

These classes exist so that ``state_dict()`` / ``load_state_dict(strict=True)`` round-trip the
reference's exact keys (``layers.<name>.<attr path>``; attribute names follow
src/yolo/blocks/*.py of the reference because the checkpoint schema is the contract).  They hold
weights only: the arithmetic of an eval-mode call is compiled by ``engine.py`` into a flat list of
libyre launches -- nothing here calls a torch conv / cuDNN, and train mode is out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_ACTS = {"silu": nn.SiLU, "none": nn.Identity}


class _EngineModule(nn.Module):
    """forward() of every block goes through the sm_100a engine (eval mode, CUDA tensors)."""

    def forward(self, x):
        from .engine import run_module
        return run_module(self, x)


def make_act(name: str) -> nn.Module:
    if name not in _ACTS:
        raise ValueError(f"activation '{name}' is not supported by the B200 engine (silu | none)")
    return _ACTS[name]()


class Conv(_EngineModule):
    """conv (no bias) + BatchNorm(eps 1e-3, momentum 0.03) + activation; pad = k//2.
    Reference: src/yolo/blocks/conv.py:55-93."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, groups=1, dilation=1,
                 activation="silu"):
        super().__init__()
        if dilation != 1:
            raise ValueError("dilated convolutions are not on the hot path")
        pad = kernel_size // 2 if padding is None else padding
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, pad, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03)
        self.act = make_act(activation)


class RepConv(_EngineModule):
    """3x3 and 1x1 branches summed before one activation (src/yolo/blocks/conv.py:109-145); the
    engine folds the 1x1 into the 3x3 centre tap."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, groups=1, activation="silu"):
        super().__init__()
        if kernel_size != 3 or padding != 1:
            raise ValueError("RepConv is 3x3 / pad 1 only")
        self.conv1 = Conv(in_channels, out_channels, 3, stride, 1, groups, activation="none")
        self.conv2 = Conv(in_channels, out_channels, 1, stride, 0, groups, activation="none")
        self.act = make_act(activation)


class RepNBottleneck(_EngineModule):
    """x + conv2(RepConv(x)) (src/yolo/blocks/bottleneck.py:26-51)."""

    def __init__(self, in_channels, out_channels, shortcut=True, groups=1, kernel_sizes=(3, 3), expansion_ratio=0.5):
        super().__init__()
        hidden = int(out_channels * expansion_ratio)
        self.conv1 = RepConv(in_channels, hidden, kernel_sizes[0], 1)
        self.conv2 = Conv(hidden, out_channels, kernel_sizes[1], 1, groups=groups)
        self.add = shortcut and in_channels == out_channels


class RepNCSP(_EngineModule):
    """conv3(cat(bottlenecks(conv1(x)), conv2(x))) (src/yolo/blocks/csp.py:28-60)."""

    def __init__(self, in_channels, out_channels, num_repeats=1, shortcut=True, groups=1, expansion_ratio=0.5):
        super().__init__()
        hidden = int(out_channels * expansion_ratio)
        self.conv1 = Conv(in_channels, hidden, 1, 1)
        self.conv2 = Conv(in_channels, hidden, 1, 1)
        self.conv3 = Conv(2 * hidden, out_channels, 1)
        self.bottlenecks = nn.Sequential(*[RepNBottleneck(hidden, hidden, shortcut, groups, expansion_ratio=1.0)
                                           for _ in range(num_repeats)])


class RepNCSPELAN4(_EngineModule):
    """src/yolo/blocks/gelan.py:27-62."""

    def __init__(self, in_channels, out_channels, hidden_channels, block_channels, num_repeats=1):
        super().__init__()
        self.conv_in = Conv(in_channels, hidden_channels, 1, 1)
        self.block1 = nn.Sequential(RepNCSP(hidden_channels // 2, block_channels, num_repeats),
                                    Conv(block_channels, block_channels, 3, 1))
        self.block2 = nn.Sequential(RepNCSP(block_channels, block_channels, num_repeats),
                                    Conv(block_channels, block_channels, 3, 1))
        self.conv_out = Conv(hidden_channels + 2 * block_channels, out_channels, 1, 1)


class ADown(_EngineModule):
    """src/yolo/blocks/downsample.py:24-46."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv_stride = Conv(in_channels // 2, out_channels // 2, 3, 2, 1)
        self.conv_pool = Conv(in_channels // 2, out_channels // 2, 1, 1, 0)


class SPPELAN(_EngineModule):
    """src/yolo/blocks/sppelan.py:24-48 (the three MaxPool2d(5,1,2) own no parameters)."""

    def __init__(self, in_channels, out_channels, hidden_channels):
        super().__init__()
        self.conv_in = Conv(in_channels, hidden_channels, 1, 1)
        self.conv_out = Conv(4 * hidden_channels, out_channels, 1, 1)


class Concat(_EngineModule):
    """src/yolo/blocks/common.py:20-37 (channel concat only)."""

    def __init__(self, dimension=1):
        super().__init__()
        if dimension != 1:
            raise ValueError("only channel concatenation is supported")
        self.dimension = dimension


class Silence(_EngineModule):
    """src/yolo/blocks/common.py:40-50."""


class Upsample(_EngineModule):
    """nn.Upsample(scale_factor=2, mode='nearest') as built at src/yolo/model/parser.py:159-171."""

    def __init__(self, scale_factor=2, mode="nearest"):
        super().__init__()
        if scale_factor != 2 or mode != "nearest":
            raise ValueError("only nearest x2 upsampling is on the hot path")
        self.scale_factor, self.mode = scale_factor, mode


class CBLinear(_EngineModule):
    """1x1 conv with bias, output split into chunks (src/yolo/blocks/auxiliary.py:30-66)."""

    def __init__(self, in_channels, out_channels_list, kernel_size=1, stride=1, padding=None, groups=1):
        super().__init__()
        if kernel_size != 1 or stride != 1 or groups != 1:
            raise ValueError("CBLinear is a plain 1x1 projection on the hot path")
        self.out_channels_list = list(out_channels_list)
        self.conv = nn.Conv2d(in_channels, sum(out_channels_list), 1, 1, 0, bias=True)


class CBFuse(_EngineModule):
    """src/yolo/blocks/auxiliary.py:76-114."""

    def __init__(self, idx):
        super().__init__()
        self.idx = list(idx)
