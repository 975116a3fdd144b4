"""Feature pyramid neck and the encoder-only feature extractor of the downstream detectors (BASELINE config 5).

`FPN` mirrors nerf_rpn/model/fpn.py:8-185 (constructor signature, `lateral_convs.i.*` / `fpn_convs.i.*` state-dict keys with
Conv3d-shaped weights, xavier init); `SwinTransformer_FPN_Pretrained_Skip` mirrors nerf_rpn/model/feature_extractor.py:1067-1187
(`base` = the MAE model without its decoder, `fpn_neck`).  The arithmetic runs in libnmae.so: 1x1x1 laterals on the tcgen05
linear kernels, the top-down nearest-neighbour adds in one kernel each, the 3x3x3 output convolutions on the tcgen05
implicit-GEMM kernel with the 256 channels zero-padded to 288 (= 6 x 48) inside the operand image.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import Tensor, nn

from . import functional as NF
from ._lib import call, conv3_image_bytes
from .swin_mae3d import SWIN_CONFIGS, SwinTransformer_MAE3D_New
from .unetr_block import from_channels_last, to_channels_last


def conv3x3x3_cl(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """nn.Conv3d(k=3, padding=1) on a channels-last volume.  Without autograd and with a channel count that is a multiple of 8
    the tensor-core kernel is used even when the count is not a multiple of 48: the operand image is built with the next multiple
    of 48 channels (zeros past the end) and the weights are zero-padded to match."""
    B, X, Y, Z, Cin = x.shape
    Co = weight.shape[0]
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad))
    Cp = (Cin + 47) // 48 * 48
    # "fp16" precision mode: channel groups of 48 or 64 run on the single-pass tcgen05 kernels with or without autograd
    h_ok = NF.get_conv_precision() == "fp16" and (Cin % 48 == 0 or Cin % 64 == 0) and (Cin % 48 == 0) == (Co % 48 == 0) and \
        (Co % 48 == 0 or Co % 64 == 0)
    if h_ok or needs_grad or Cin % 48 == 0 or Cin % 8 != 0 or Co % 16 != 0:
        return NF.Conv3x3x3Fn.apply(x, weight, bias)
    x = x.contiguous().float()
    wp = torch.zeros(Co, Cp, 3, 3, 3, device=x.device, dtype=torch.float32)
    wp[:, :Cin] = weight
    img = torch.empty(conv3_image_bytes(B, X, Y, Z, Cp), dtype=torch.uint8, device=x.device)
    call("nmae_conv3_image_build", x, Cin, 0, B, X, Y, Z, Cp, 0, img, device=x.device)
    y = torch.empty(B, X, Y, Z, Co, device=x.device, dtype=torch.float32)
    wws = torch.empty(27 * Cp * Co, device=x.device, dtype=torch.float32)
    call("nmae_conv3x3x3_fwd", None, img, wp, None if bias is None else bias.contiguous().float(), B, X, Y, Z, Cp, Co, wws, y,
         device=x.device)
    return y


class FPN(nn.Module):
    """nerf_rpn/model/fpn.py:8-185.  Inputs/outputs are (B,C,X,Y,Z) tensors like the reference's; `forward_cl` takes and returns
    channels-last (B,X,Y,Z,C) tensors and avoids the layout copies (the encoder produces channels-last natively)."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 extra_convs_on_inputs=True, relu_before_extra_convs=False, upsample_cfg=dict(mode="nearest")):
        super().__init__()
        assert isinstance(in_channels, list)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.upsample_cfg = dict(upsample_cfg)
        if self.upsample_cfg != dict(mode="nearest"):
            raise NotImplementedError("FPN: only the reference default upsample_cfg=dict(mode='nearest') is implemented")
        if add_extra_convs:
            raise NotImplementedError("FPN: add_extra_convs is not used by the NeRF-MAE feature extractors (fpn.py:119-129)")
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels)
            assert num_outs == end_level - start_level
        self.start_level = start_level
        self.end_level = end_level
        self.add_extra_convs = add_extra_convs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(nn.Conv3d(in_channels[i], out_channels, 1))       # parameter holders: fpn.py:109-110
            self.fpn_convs.append(nn.Conv3d(out_channels, out_channels, 3, padding=1))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward_cl(self, inputs: Sequence[Tensor]) -> List[Tensor]:
        assert len(inputs) == len(self.in_channels)
        laterals = []
        for i, conv in enumerate(self.lateral_convs):                                     # fpn.py:139-142
            x = inputs[i + self.start_level]
            laterals.append(NF.LinearFn.apply(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias))
        for i in range(len(laterals) - 1, 0, -1):                                         # fpn.py:146-158
            fine, coarse = laterals[i - 1], laterals[i]
            if torch.is_grad_enabled() and (fine.requires_grad or coarse.requires_grad):
                idx = [torch.div(torch.arange(fine.shape[d], device=fine.device) * coarse.shape[d], fine.shape[d], rounding_mode="floor")
                       for d in (1, 2, 3)]
                laterals[i - 1] = fine + coarse[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]]
            else:
                B, Xf, Yf, Zf, C = fine.shape
                call("nmae_upsample_nearest_add", fine, coarse, B, Xf, Yf, Zf, coarse.shape[1], coarse.shape[2], coarse.shape[3], C,
                     device=fine.device)
        outs = [conv3x3x3_cl(laterals[i], self.fpn_convs[i].weight, self.fpn_convs[i].bias) for i in range(len(laterals))]
        for _ in range(self.num_outs - len(outs)):                                        # fpn.py:167-170: max_pool3d(k=1, stride=2)
            outs.append(outs[-1][:, ::2, ::2, ::2].contiguous())
        return outs

    def forward(self, inputs):
        return tuple(from_channels_last(o) for o in self.forward_cl([to_channels_last(x) for x in inputs]))


class SwinTransformer_FPN_Pretrained_Skip(nn.Module):
    """nerf_rpn/model/feature_extractor.py:1067-1187: the pretrained MAE encoder (decoder deleted) + FPN neck.  The reference
    hard-codes swin_s; `backbone_type` selects the other widths of run_swin_mae3d.py:378-399 (BASELINE config 5 uses swin_l)."""

    def __init__(self, expand_dim: bool = True, out_channels: int = 256, resolution=160, checkpoint_path=None, is_eval=False,
                 backbone_type: str = "swin_s"):
        super().__init__()
        self.out_channels = out_channels
        cfg = SWIN_CONFIGS[backbone_type]
        model = SwinTransformer_MAE3D_New(patch_size=[4, 4, 4], embed_dim=cfg["embed_dim"], depths=cfg["depths"],
                                          num_heads=cfg["num_heads"], window_size=[4, 4, 4], stochastic_depth_prob=0.1,
                                          expand_dim=True, resolution=resolution)
        if not is_eval:
            if checkpoint_path is None:
                raise AssertionError("The checkpoint does not exist.")
            checkpoint = torch.load(checkpoint_path, map_location="cpu")
            model.load_state_dict(checkpoint["state_dict"])
        del model.decoder4, model.decoder3, model.decoder2, model.decoder1, model.out, model.mask_token
        fpn_in = [cfg["embed_dim"] * 2 ** i if expand_dim else cfg["embed_dim"] for i in range(len(cfg["depths"]))]
        self.base = model
        self.fpn_neck = FPN(fpn_in, out_channels, len(fpn_in))

    def forward_features_cl(self, x: Tensor) -> List[Tensor]:  # noqa: D401 (shared with SwinTransformer_FPN_Pretrained below)
        """(B,4,R,R,R) -> the four stage outputs, channels-last (feature_extractor.py:1171-1184 without the permute copies)."""
        t = self.base.patch_partition(x, self.base.pos_embed.view(-1, self.base.embed_dim))
        feats = []
        for stage in self.base.stages:
            t = stage(t)
            feats.append(t)
        return feats

    def forward(self, x: Tensor):
        outs = self.fpn_neck.forward_cl(self.forward_features_cl(x))
        return tuple(from_channels_last(o) for o in outs)


class SwinTransformer_FPN_Pretrained(SwinTransformer_FPN_Pretrained_Skip):
    """nerf_rpn/model/feature_extractor.py:1190-1307: the same extractor built on the LEGACY `SwinTransformer_MAE3D` (conv +
    trilinear-upsample decoder in its checkpoints, deleted after loading together with the mask token)."""

    def __init__(self, expand_dim: bool = True, out_channels: int = 256, resolution=160, checkpoint_path=None, is_eval=False,
                 backbone_type: str = "swin_s"):
        nn.Module.__init__(self)
        from .swin_mae3d_legacy import SwinTransformer_MAE3D
        self.out_channels = out_channels
        cfg = SWIN_CONFIGS[backbone_type]
        model = SwinTransformer_MAE3D(patch_size=[4, 4, 4], embed_dim=cfg["embed_dim"], depths=cfg["depths"], num_heads=cfg["num_heads"],
                                      window_size=[4, 4, 4], stochastic_depth_prob=0.1, expand_dim=True, resolution=resolution)
        if not is_eval:
            if checkpoint_path is None:
                raise AssertionError("The checkpoint does not exist.")
            checkpoint = torch.load(checkpoint_path, map_location="cpu")
            model.load_state_dict(checkpoint["state_dict"])
        del model.decoder_layers, model.mask_token
        fpn_in = [cfg["embed_dim"] * 2 ** i if expand_dim else cfg["embed_dim"] for i in range(len(cfg["depths"]))]
        self.base = model
        self.fpn_neck = FPN(fpn_in, out_channels, len(fpn_in))
