"""Decoder blocks with the reference's names, constructor signatures and state-dict keys
(nerf_mae/model/mae/unetr_block.py), computing through libnmae.so on channels-last volumes.

The nn.Conv3d / nn.ConvTranspose3d members only hold the parameters (same shapes, same default
initialisation, same RNG consumption order as the reference); their own forward is never called.
Module boundaries speak (B,C,X,Y,Z) like the reference; internally memory is (B,X,Y,Z,C), so tensors
handed from one of these blocks to the next are never re-laid-out.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
import torch.nn as nn

from . import functional as NF


def to_channels_last(x: torch.Tensor) -> torch.Tensor:
    """(B,C,X,Y,Z) with any strides -> contiguous (B,X,Y,Z,C); free when x already is channels-last memory."""
    return x.permute(0, 2, 3, 4, 1).contiguous()


def from_channels_last(y: torch.Tensor) -> torch.Tensor:
    return y.permute(0, 4, 1, 2, 3)


class UnetResBlock(nn.Module):
    """unetr_block.py:23-93 (instance norm + LeakyReLU(0.01) only, which is all the MAE path uses)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, norm_name="instancenorm", act_name="leakyrelu",
                 dropout=None):
        super().__init__()
        if kernel_size != 3 or stride != 1:
            raise ValueError("UnetResBlock: the MAE decoder uses kernel_size=3, stride=1")
        if norm_name != "instancenorm":
            raise ValueError(f"Unsupported normalization: {norm_name}")
        if act_name != "leakyrelu":
            raise ValueError(f"Unsupported activation: {act_name}")
        if dropout is not None:
            raise ValueError("UnetResBlock: dropout is not used by the MAE decoder")
        self.conv1 = nn.Conv3d(in_channels, out_channels, kernel_size, stride, padding=kernel_size // 2)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2)
        self.norm1 = nn.InstanceNorm3d(out_channels)
        self.norm2 = nn.InstanceNorm3d(out_channels)
        self.activation = nn.LeakyReLU(0.01, inplace=True)
        self.dropout = None
        self.downsample = in_channels != out_channels
        if self.downsample:
            self.conv3 = nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=stride)
            self.norm3 = nn.InstanceNorm3d(out_channels)

    def can_fuse_out(self, x_cl: torch.Tensor, out_block) -> bool:
        """The 1x1x1 output convolution can ride on this block (functional.ResBlockFn) when the block runs on the single-pass
        fp16 convolution path and the output block is the 4-channel one of the MAE model."""
        Cin, Co = x_cl.shape[-1], self.conv1.weight.shape[0]
        return (out_block is not None and NF.get_conv_precision() == "fp16" and out_block.conv.weight.shape[0] == 4 and Co <= 128
                and (Cin % 48 == 0 or Cin % 64 == 0) and (Co % 48 == 0 or Co % 64 == 0) and (Cin % 48 == 0) == (Co % 48 == 0))

    def forward_cl(self, x_cl: torch.Tensor, out_block=None) -> torch.Tensor:
        """out_block: a UnetOutBlock to evaluate on the block's result inside the same autograd node (see can_fuse_out)."""
        w3 = b3 = None
        if self.downsample:
            w3 = self.conv3.weight.view(self.conv3.weight.shape[0], -1)
            b3 = self.conv3.bias
        if out_block is not None:
            wo = out_block.conv.weight.view(out_block.conv.weight.shape[0], -1)
            return NF.ResBlockFn.apply(x_cl, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, w3, b3,
                                       float(self.activation.negative_slope), wo, out_block.conv.bias)
        return NF.ResBlockFn.apply(x_cl, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, w3, b3,
                                   float(self.activation.negative_slope))

    def forward(self, x):
        return from_channels_last(self.forward_cl(to_channels_last(x)))


class UnetOutBlock(nn.Module):
    """unetr_block.py:96-116: 1x1x1 conv."""

    def __init__(self, in_channels: int, out_channels: int, dropout: Optional[float] = None):
        super().__init__()
        if dropout is not None:
            raise ValueError("UnetOutBlock: dropout is not used by the MAE decoder")
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=1, bias=True)
        self.dropout = None

    def forward_cl(self, x_cl):
        return NF.linear(x_cl, self.conv.weight.view(self.conv.weight.shape[0], -1), self.conv.bias)

    def forward(self, inp):
        return from_channels_last(self.forward_cl(to_channels_last(inp)))


class UnetrUpBlock(nn.Module):
    """unetr_block.py:119-200: ConvTranspose3d(kernel == stride) -> cat(skip) -> UnetResBlock."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[Sequence[int], int],
                 upsample_kernel_size: Union[Sequence[int], int], norm_name: str = "instancenorm", res_block: bool = False,
                 use_skip: bool = True, input_padding=0) -> None:
        super().__init__()
        if not res_block:
            raise ValueError("UnetrUpBlock: only res_block=True exists in the reference (UnetBasicBlock is undefined there)")
        if input_padding != 0 or not isinstance(upsample_kernel_size, int):
            raise ValueError("UnetrUpBlock: integer upsample kernel (== stride) without padding only")
        self.use_skip = use_skip
        self.transp_conv = nn.ConvTranspose3d(in_channels, out_channels, upsample_kernel_size, stride=upsample_kernel_size,
                                              padding=input_padding, output_padding=0)
        self.conv_block = UnetResBlock(out_channels + out_channels if use_skip else out_channels, out_channels,
                                       kernel_size=kernel_size, stride=1, norm_name=norm_name)

    def forward_cl(self, inp_cl, skip_cl=None, out_block=None):
        """out_block: optional UnetOutBlock applied to the result (fused into the residual block's autograd node when possible)."""
        k = self.transp_conv.kernel_size[0]
        if self.use_skip and skip_cl is None:
            raise ValueError("UnetrUpBlock(use_skip=True) needs a skip tensor")
        up = NF.ConvTransposeCatFn.apply(inp_cl, self.transp_conv.weight, self.transp_conv.bias,
                                         skip_cl if self.use_skip else None, k)
        if out_block is None:
            return self.conv_block.forward_cl(up)
        if self.conv_block.can_fuse_out(up, out_block):
            return self.conv_block.forward_cl(up, out_block)
        return out_block.forward_cl(self.conv_block.forward_cl(up))

    def forward(self, inp, skip=None):
        skip_cl = to_channels_last(skip) if (self.use_skip and skip is not None) else None
        return from_channels_last(self.forward_cl(to_channels_last(inp), skip_cl))
