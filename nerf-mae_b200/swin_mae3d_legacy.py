"""The older MAE model of the reference, `SwinTransformer_MAE3D` (nerf_mae/model/mae/swin_mae3d.py:417-1064), still used by
`inference*.py` and by `SwinTransformer_FPN_Pretrained` (nerf_rpn/model/feature_extractor.py:1190-1307) as an ENCODER whose
checkpoints carry a conv + InstanceNorm + LeakyReLU(0.2) + trilinear-upsample decoder (`decoder_layers.{0,4,8,12}.*`).

What is reproduced: the constructor signature and creation order (bit-identical initial weights under the same torch seed), the
state-dict keys, `patch_partition` / `pos_embed` / `stages` / `mask_token`, `window_masking_3d` with the three sampling
strategies ("random", "grid", "block" - the latter consumes numpy's global RNG exactly like the reference), `transform`,
`forward_encoder`, `forward_decoder`, `patchify_3d`, `forward_loss`, and the 7-tuple eval return of `forward`.

What the reference itself cannot do: its `forward()` always fails - `forward_loss` calls `patchify_3d(pred)` on the decoder output
`(N,40,40,40,out_channels)`, which trips `assert x.shape[2] == x.shape[3] == x.shape[4]` (swin_mae3d.py:838,927) [verified on the
live reference].  `forward()` here raises the same AssertionError for such outputs, so callers see the reference's behaviour; the
encoder / decoder halves, which is what the reference's users actually call, run on libnmae.so.
"""
from __future__ import annotations

import random
from functools import partial
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn
from torchvision.ops.misc import Permute

from . import functional as NF
from .swin_mae3d import LayerNorm, PatchMerging, PatchPartition, SwinTransformerBlock, draw_block_mask
from .torch_utils import get_3d_sincos_pos_embed


def draw_legacy_mask(n_tok, p_remove: float, strategy: str, block: int = 4) -> np.ndarray:
    """Token mask (H,W,D) uint8 of `window_masking_3d` (swin_mae3d.py:630-773) for the three strategies, drawn on the host with the
    reference's RNG consumption: "random" one `random.random()` per block (h-major), "grid" none (three of every four blocks in
    h-major order), "block" three `np.random.randint` draws."""
    H, W, D = n_tok
    nh, nw, nd = H // block, W // block, D // block
    if strategy == "random":
        return draw_block_mask(n_tok, p_remove, block)
    m = np.zeros((H, W, D), dtype=np.uint8)
    if strategy not in ("grid", "block"):
        return m                  # the reference's if/elif chain matches nothing (e.g. the constructor default None): no masking
    idx = [(h, w, d) for h in range(nh) for w in range(nw) for d in range(nd)]
    if strategy == "grid":
        count = 0
        for h, w, d in idx:
            if count in (0, 1, 2):
                m[h * block:(h + 1) * block, w * block:(w + 1) * block, d * block:(d + 1) * block] = 1
                count += 1
            else:
                count += 1
                if count == 4:
                    count = 0
        return m
    if strategy == "block":
        num_to_keep = (nh * nw * nd) // 4
        for _ in range(3):
            masked = 0
            h_start = np.random.randint(0, nh - 0.25 * nh)
            for h, w, d in idx:
                if h > h_start:
                    blk = m[h * block:(h + 1) * block, w * block:(w + 1) * block, d * block:(d + 1) * block]
                    if blk.sum() == 0:
                        blk[...] = 1
                        masked += 1
                if masked >= num_to_keep:
                    break
    return m


class SwinTransformer_MAE3D(nn.Module):
    """swin_mae3d.py:417-1064."""

    def __init__(self, patch_size: List[int], embed_dim: int, depths: List[int], num_heads: List[int], window_size: List[int],
                 mlp_ratio: float = 4.0, dropout: float = 0.0, attention_dropout: float = 0.0, stochastic_depth_prob: float = 0.1,
                 norm_layer: Optional[Callable[..., nn.Module]] = partial(LayerNorm, eps=1e-5),
                 block: Optional[Callable[..., nn.Module]] = SwinTransformerBlock,
                 downsample_layer: Callable[..., nn.Module] = PatchMerging, expand_dim: bool = True, out_channels: int = 256,
                 input_dim: int = 4, decoder_embed_dim: int = 768, masking_prob=0.50, resolution=160, masking_strategy=None):
        super().__init__()
        if input_dim != 4 or len(set(patch_size)) != 1:
            raise ValueError("RGB+sigma grids with cubic patches only")
        self.out_channels = out_channels
        self.sampling_strategy = masking_strategy
        self.patch_size = patch_size
        self.masking_prob = masking_prob
        self.resolution = resolution
        self.embed_dim = embed_dim
        self.patch_partition = PatchPartition(
            nn.Conv3d(input_dim, embed_dim, kernel_size=tuple(patch_size), stride=tuple(patch_size)),
            Permute([0, 2, 3, 4, 1]),
            norm_layer(embed_dim),
        )
        self.stages = nn.ModuleList()
        total_stage_blocks = sum(depths)
        stage_block_id = 0
        dims = []
        for i_stage in range(len(depths)):
            stage = nn.ModuleList()
            dim = embed_dim * 2 ** i_stage if expand_dim else embed_dim
            dims.append(dim)
            if i_stage > 0:
                stage.append(downsample_layer(dims[-2], norm_layer, expand_dim))
            for i_layer in range(depths[i_stage]):
                sd_prob = stochastic_depth_prob * float(stage_block_id) / (total_stage_blocks - 1)
                stage.append(block(dim, num_heads[i_stage], window_size=window_size,
                                   shift_size=[0 if i_layer % 2 == 0 else w // 2 for w in window_size], mlp_ratio=mlp_ratio,
                                   dropout=dropout, attention_dropout=attention_dropout, stochastic_depth_prob=sd_prob,
                                   norm_layer=norm_layer))
                stage_block_id += 1
            self.stages.append(nn.Sequential(*stage))
        self.num_patches = int(round(self.resolution // patch_size[0]))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, self.num_patches, self.num_patches, embed_dim),
                                      requires_grad=False)
        self.mask_token = nn.Parameter(torch.zeros(embed_dim))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        size = (40, 40, 40)                                           # hard-coded in the reference (swin_mae3d.py:590)
        # parameter holders with the reference's Sequential indices (state-dict keys decoder_layers.{0,4,8,12}.{weight,bias})
        self.decoder_layers = nn.Sequential(
            nn.Conv3d(decoder_embed_dim, 512, kernel_size=3, stride=1, padding=1), nn.InstanceNorm3d(512), nn.LeakyReLU(0.2, inplace=True),
            nn.Upsample(size=(10, 10, 10), mode="trilinear", align_corners=False),
            nn.Conv3d(512, 256, kernel_size=3, stride=1, padding=1), nn.InstanceNorm3d(256), nn.LeakyReLU(0.2, inplace=True),
            nn.Upsample(size=(20, 20, 20), mode="trilinear", align_corners=False),
            nn.Conv3d(256, 128, kernel_size=3, stride=1, padding=1), nn.InstanceNorm3d(128), nn.LeakyReLU(0.2, inplace=True),
            nn.Upsample(size=size, mode="trilinear", align_corners=False),
            nn.Conv3d(128, out_channels, kernel_size=3, stride=1, padding=1),
        )
        self.alpha_activation = nn.Sigmoid()
        self.initialize_weights()
        self._tok_mask_u8 = None

    def initialize_weights(self):
        pos_embed = get_3d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.num_patches), cls_token=False)
        self.pos_embed.data.copy_(torch.from_numpy(pos_embed).float())
        torch.nn.init.normal_(self.mask_token, std=0.02)

    # ------------------------------------------------------------------ masking (swin_mae3d.py:630-773)
    def window_masking_3d(self, x, patch_size=(4, 4, 4), p_remove=0.50, mask_token=None, sampling_strategy="random"):
        B, H, W, D, C = x.shape
        m = torch.from_numpy(draw_legacy_mask((H, W, D), p_remove, sampling_strategy, patch_size[0])).to(x.device)
        mb = m.bool()[None, ..., None]
        fill = torch.zeros(C, device=x.device, dtype=x.dtype) if mask_token is None else mask_token.to(x.device)
        return torch.where(mb, fill.view(1, 1, 1, 1, C), x), mb.expand(B, H, W, D, 1).to(x.dtype)

    # ------------------------------------------------------------------ helpers
    def patchify_3d(self, x, mask=None):
        """swin_mae3d.py:829-852 (including its shape assertion)."""
        p = self.patch_size[0]
        assert x.shape[2] == x.shape[3] == x.shape[4] and x.shape[2] % p == 0
        n = x.shape[2] // p
        out = x.reshape(x.shape[0], 4, n, p, n, p, n, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(x.shape[0], n, n, n, p ** 3, 4)
        if mask is not None:
            m = mask.reshape(x.shape[0], 4, n, p, n, p, n, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(x.shape[0], n, n, n, p ** 3, 4)
            return out, m[..., 0].unsqueeze(-1).int()
        return out

    def transform(self, x):
        """swin_mae3d.py:880-896: list of (4,X,Y,Z) -> (B,4,R,R,R) batch + (B,3) extents (the pad mask, compactly)."""
        return NF.pad_grids(x, self.resolution)

    def forward_encoder(self, x):
        """swin_mae3d.py:898-916: patch embed + pos + masking (strategy of the constructor) + the four stages.
        x (B,4,R,R,R) -> (latent (B,h,w,d,8C) channels-last, mask_patches (B,n,n,n,1) float)."""
        n = self.num_patches
        B = x.shape[0]
        m_np = draw_legacy_mask((n, n, n), self.masking_prob, self.sampling_strategy, 4)
        tok_mask = torch.from_numpy(m_np).to(x.device, non_blocking=True)
        self._tok_mask_u8 = tok_mask
        t = self.patch_partition(x, self.pos_embed.view(-1, self.embed_dim), tok_mask.view(-1), self.mask_token)
        for stage in self.stages:
            t = stage(t)
        return t, tok_mask.view(1, n, n, n, 1).expand(B, n, n, n, 1).to(x.dtype)

    def forward_decoder(self, latent):
        """swin_mae3d.py:1026-1034: (B,h,w,d,C) channels-last -> (B,40,40,40,out_channels) channels-last.  conv3^3 on the tensor
        cores, InstanceNorm + LeakyReLU(0.2) and the trilinear upsampling in libnmae.so."""
        layers = self.decoder_layers
        y = latent
        for i in (0, 4, 8):
            conv, up = layers[i], layers[i + 3]
            y = NF.Conv3x3x3Fn.apply(y, conv.weight, conv.bias)
            y = NF.InstNormLReLUFn.apply(y, layers[i + 2].negative_slope, layers[i + 1].eps)
            y = NF.UpsampleTrilinearFn.apply(y, up.size)
        return NF.Conv3x3x3Fn.apply(y, layers[12].weight, layers[12].bias)

    def forward_loss(self, x, pred, ext, mask_patches, is_eval=False):
        """swin_mae3d.py:924-975.  As in the reference, `pred` must be an (N,4,R,R,R)-shaped volume for `patchify_3d`; the decoder
        of this class produces (N,40,40,40,out_channels), for which the reference asserts."""
        assert pred.dim() == 5 and pred.shape[2] == pred.shape[3] == pred.shape[4] and pred.shape[2] % self.patch_size[0] == 0, \
            "patchify_3d(pred): the reference asserts on this decoder's output shape (swin_mae3d.py:838)"
        tok = (mask_patches[0, ..., 0] != 0).to(torch.uint8).contiguous()
        out3 = NF.MAELossFn.apply(pred.permute(0, 2, 3, 4, 1).contiguous(), x, ext, tok, self.patch_size[0])
        if not is_eval:
            return out3[0], out3[1], out3[2]
        target = self.patchify_3d(x)
        return out3[0], out3[1], out3[2], self.patchify_3d(pred), target[..., 3].unsqueeze(-1) > 0.01, target

    def forward(self, x, is_eval=False):
        """swin_mae3d.py:1036-1064.  Raises the reference's AssertionError (see the module docstring) after running the encoder
        and the decoder, exactly where the reference does."""
        xb, ext = self.transform(x)
        latent, mask_patches = self.forward_encoder(xb)
        pred = self.forward_decoder(latent)
        if is_eval:
            loss, loss_rgb, loss_alpha, pred_rgb, mask, target_rgb = self.forward_loss(xb, pred, ext, mask_patches, True)
            return loss, loss_rgb, loss_alpha, pred_rgb, mask, mask_patches, target_rgb
        return self.forward_loss(xb, pred, ext, mask_patches, False)
