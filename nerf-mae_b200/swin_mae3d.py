"""B200-native drop-in for nerf_mae/model/mae/swin_mae3d.py (the `_New` MAE model and its Swin blocks).

Same class names, constructor signatures, attribute names and state-dict keys as the reference
(SURVEY.md A.4), so reference checkpoints load unchanged and callers that reach into the model
(`patch_partition`, `pos_embed`, `stages[i]`, `decoder4..1`, `out`, `mask_token`) keep working.  The torch
modules held inside (nn.Linear, nn.Conv3d, torchvision MLP ...) are parameter containers created in the
reference's order - a model built under the same torch seed has bit-identical initial weights - while all
arithmetic runs in libnmae.so (functional.py).  There is no PyTorch fallback.
"""
from __future__ import annotations

import random
from functools import partial
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor
from torchvision.ops.misc import MLP, Permute
from torchvision.ops.stochastic_depth import StochasticDepth

from . import functional as NF
from .torch_utils import get_3d_sincos_pos_embed
from .unetr_block import UnetOutBlock, UnetrUpBlock, from_channels_last, to_channels_last


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm whose forward runs the nmae kernel (state-dict compatible)."""

    def forward(self, x: Tensor) -> Tensor:
        if len(self.normalized_shape) != 1 or self.weight is None or self.bias is None:
            raise ValueError("nmae LayerNorm: affine LayerNorm over the last dim only")
        return NF.layer_norm(x, self.weight, self.bias, self.eps)


def _check_window(window_size, shift_size):
    if len(window_size) != 3 or len(shift_size) != 3:
        raise ValueError("window_size and shift_size must be of length 3")  # swin_mae3d.py:231-232
    if list(window_size) != [4, 4, 4]:
        raise ValueError("the nmae W-MSA kernel is specialised for the 4x4x4 windows every reference config uses")
    if len(set(shift_size)) != 1 or shift_size[0] not in (0, 2):
        raise ValueError("shift_size must be [0,0,0] or [2,2,2] (window//2), as built by the reference model")


def shifted_window_attention(input: Tensor, qkv_weight: Tensor, proj_weight: Tensor, relative_position_bias: Tensor,
                             window_size: List[int], num_heads: int, shift_size: List[int], attention_dropout: float = 0.0,
                             dropout: float = 0.0, qkv_bias: Optional[Tensor] = None, proj_bias: Optional[Tensor] = None,
                             logit_scale: Optional[Tensor] = None, *, relative_position_bias_table: Optional[Tensor] = None):
    """Functional W-MSA with the reference signature (swin_mae3d.py:27-40).

    The kernel consumes the (343, nH) table directly; when only the gathered (1,nH,64,64) bias is given it is
    folded back to a table (exact: every table entry appears in the gathered bias)."""
    _check_window(window_size, shift_size)
    if attention_dropout != 0.0 or dropout != 0.0:
        raise ValueError("dropout is 0 in every reference config; not implemented")
    if logit_scale is not None:
        raise ValueError("logit_scale (Swin-V2) is not used by the reference MAE path")
    table = relative_position_bias_table
    if table is None:
        idx = _relative_position_index(window_size).to(relative_position_bias.device)
        flat = relative_position_bias.reshape(num_heads, -1).t()          # (4096, nH)
        table = torch.zeros(343, num_heads, device=flat.device, dtype=flat.dtype).index_copy(0, idx, flat)
    return NF.window_attention(input, qkv_weight, qkv_bias, proj_weight, proj_bias, table, num_heads, int(shift_size[0]))


def _relative_position_index(window_size) -> Tensor:
    """swin_mae3d.py:257-280: ((dh+3)*7 + (dw+3))*7 + (dd+3), tokens flattened h-major."""
    ws = window_size
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 2] += ws[2] - 1
    rel[:, :, 0] *= (2 * ws[2] - 1) * (2 * ws[1] - 1)
    rel[:, :, 1] *= 2 * ws[2] - 1
    return rel.sum(-1).flatten()


class ShiftedWindowAttention(nn.Module):
    """swin_mae3d.py:214-307."""

    def __init__(self, dim: int, window_size: List[int], shift_size: List[int], num_heads: int, qkv_bias: bool = True,
                 proj_bias: bool = True, attention_dropout: float = 0.0, dropout: float = 0.0):
        super().__init__()
        _check_window(window_size, shift_size)
        if dim != 32 * num_heads:
            raise ValueError(f"head_dim must be 32 (dim={dim}, heads={num_heads}): every runnable reference config has it")
        if attention_dropout != 0.0 or dropout != 0.0:
            raise ValueError("dropout is 0 in every reference config; not implemented")
        self.window_size = window_size
        self.shift_size = shift_size
        self.num_heads = num_heads
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1), num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.register_buffer("relative_position_index", _relative_position_index(window_size))

    def get_relative_position_bias(self) -> Tensor:
        N = 64
        b = self.relative_position_bias_table[self.relative_position_index].view(N, N, -1)
        return b.permute(2, 0, 1).contiguous().unsqueeze(0)

    def forward(self, x: Tensor, *, ln: Optional[nn.LayerNorm] = None, residual: bool = False,
                row_scale: Optional[Tensor] = None) -> Tensor:
        """x: [B,H,W,D,C] -> same.  `ln`/`residual`/`row_scale` let SwinTransformerBlock fuse norm1, the
        residual add and the stochastic-depth scale into the same call."""
        return NF.window_attention(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                                   self.relative_position_bias_table, self.num_heads, int(self.shift_size[0]),
                                   ln_w=None if ln is None else ln.weight, ln_b=None if ln is None else ln.bias,
                                   eps=1e-5 if ln is None else ln.eps, residual=residual, row_scale=row_scale)


class SwinTransformerBlock(nn.Module):
    """swin_mae3d.py:310-369."""

    def __init__(self, dim: int, num_heads: int, window_size: List[int], shift_size: List[int], mlp_ratio: float = 4.0,
                 dropout: float = 0.0, attention_dropout: float = 0.0, stochastic_depth_prob: float = 0.0,
                 norm_layer: Callable[..., nn.Module] = LayerNorm, attn_layer: Callable[..., nn.Module] = ShiftedWindowAttention):
        super().__init__()
        if dropout != 0.0:
            raise ValueError("dropout is 0 in every reference config; not implemented")
        self.norm1 = norm_layer(dim)
        self.attn = attn_layer(dim, window_size, shift_size, num_heads, attention_dropout=attention_dropout, dropout=dropout)
        self.stochastic_depth = StochasticDepth(stochastic_depth_prob, "row")
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(dim, [int(dim * mlp_ratio), dim], activation_layer=nn.GELU, inplace=None, dropout=dropout)
        for m in self.mlp.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, std=1e-6)

    def _sd_scale(self, x: Tensor) -> Optional[Tensor]:
        """torchvision.ops.stochastic_depth (row mode): identical RNG consumption, returned as a (B,) scale."""
        p = self.stochastic_depth.p
        if not self.training or p == 0.0:
            return None
        survival = 1.0 - p
        noise = torch.empty([x.shape[0]] + [1] * (x.ndim - 1), dtype=x.dtype, device=x.device).bernoulli_(survival)
        if survival > 0.0:
            noise.div_(survival)
        return noise.view(-1)

    def forward(self, x: Tensor):
        fused = isinstance(self.attn, ShiftedWindowAttention) and isinstance(self.norm1, nn.LayerNorm) and \
            isinstance(self.norm2, nn.LayerNorm)
        if not fused:  # foreign attn_layer / norm_layer injected through the reference's plugin points
            s = self._sd_scale(x)
            a = self.attn(self.norm1(x))
            x = x + (a if s is None else a * s.view(-1, 1, 1, 1, 1))
            s = self._sd_scale(x)
            m = NF.mlp(self.norm2(x), self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias)
            return x + (m if s is None else m * s.view(-1, 1, 1, 1, 1))
        x = self.attn(x, ln=self.norm1, residual=True, row_scale=self._sd_scale(x))
        return NF.mlp(x, self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias, ln_w=self.norm2.weight,
                      ln_b=self.norm2.bias, eps=self.norm2.eps, residual=True, row_scale=self._sd_scale(x))


class PatchMerging(nn.Module):
    """swin_mae3d.py:372-414."""

    def __init__(self, dim: int, norm_layer: Callable[..., nn.Module] = LayerNorm, expand_dim: bool = True):
        super().__init__()
        if not expand_dim:
            raise ValueError("PatchMerging: expand_dim=False is not used by any reference config")
        self.dim = dim
        self.reduction = nn.Linear(8 * dim, dim * 2 if expand_dim else dim, bias=False)
        self.norm = norm_layer(8 * dim)

    def forward(self, x: Tensor):
        if not isinstance(self.norm, nn.LayerNorm):
            raise ValueError("PatchMerging: norm_layer must be a LayerNorm")
        return NF.PatchMergeFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps, self.reduction.weight)


class PatchPartition(nn.Sequential):
    """`patch_partition` of the reference: Sequential(Conv3d(k=s=p), Permute, LayerNorm) with the same child
    indices (keys patch_partition.0.*, patch_partition.2.*), evaluated as one fused call."""

    def forward(self, x: Tensor, pos: Optional[Tensor] = None, mask_u8: Optional[Tensor] = None,
                mask_token: Optional[Tensor] = None) -> Tensor:
        conv, norm = self[0], self[2]
        p = conv.kernel_size[0]
        return NF.PatchEmbedFn.apply(x, conv.weight, conv.bias, norm.weight, norm.bias, norm.eps, pos, mask_u8, mask_token, p)


def draw_block_mask(n_tok, p_remove: float, block: int = 4) -> np.ndarray:
    """The reference's mask draw (swin_mae3d.py:1364-1373): one Python `random.random() < p` per block^3 block of
    tokens, h-major / d-minor, consuming the global Mersenne-Twister stream exactly like the reference."""
    H, W, D = n_tok
    m = np.zeros((H, W, D), dtype=np.uint8)
    rnd = random.random
    for h in range(0, H - block + 1, block):
        for w in range(0, W - block + 1, block):
            for d in range(0, D - block + 1, block):
                if rnd() < p_remove:
                    m[h:h + block, w:w + block, d:d + block] = 1
    return m


class SwinTransformer_MAE3D_New(nn.Module):
    """swin_mae3d.py:1067-1599."""

    def __init__(self, patch_size: List[int], embed_dim: int, depths: List[int], num_heads: List[int], window_size: List[int],
                 mlp_ratio: float = 4.0, dropout: float = 0.0, attention_dropout: float = 0.0,
                 stochastic_depth_prob: float = 0.1,
                 norm_layer: Optional[Callable[..., nn.Module]] = partial(LayerNorm, eps=1e-5),
                 block: Optional[Callable[..., nn.Module]] = SwinTransformerBlock,
                 downsample_layer: Callable[..., nn.Module] = PatchMerging, expand_dim: bool = True, out_channels: int = 4,
                 input_ch_dim: int = 4, decoder_embed_dim: int = 768, masking_prob=0.50, resolution=160, drop_rate=0.10,
                 masking_strategy="random"):
        super().__init__()
        if input_ch_dim != 4 or out_channels != 4:
            raise ValueError("the MAE path works on RGB+sigma grids: input_ch_dim == out_channels == 4")
        if len(set(patch_size)) != 1:
            raise ValueError("cubic patches only")
        self.out_channels = out_channels
        self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.masking_prob = masking_prob
        self.resolution = resolution
        self.patch_partition = PatchPartition(
            nn.Conv3d(input_ch_dim, embed_dim, kernel_size=tuple(patch_size), stride=tuple(patch_size)),
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
        self.decoder4 = UnetrUpBlock(embed_dim * 8, embed_dim * 4, kernel_size=3, upsample_kernel_size=2, res_block=True)
        self.decoder3 = UnetrUpBlock(embed_dim * 4, embed_dim * 2, kernel_size=3, upsample_kernel_size=2, res_block=True)
        self.decoder2 = UnetrUpBlock(embed_dim * 2, embed_dim * 1, kernel_size=3, upsample_kernel_size=2, res_block=True)
        self.decoder1 = UnetrUpBlock(embed_dim * 1, embed_dim // 2, kernel_size=3, upsample_kernel_size=4, res_block=True,
                                     use_skip=False)
        self.out = UnetOutBlock(in_channels=embed_dim // 2, out_channels=out_channels)
        self.num_patches = int(round(self.resolution // patch_size[0]))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, self.num_patches, self.num_patches, embed_dim),
                                      requires_grad=False)
        self.mask_token = nn.Parameter(torch.zeros(embed_dim))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self.alpha_activation = nn.Sigmoid()
        self.initialize_weights()
        self._tok_mask_u8: Optional[Tensor] = None

    def initialize_weights(self):
        pos_embed = get_3d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.num_patches), cls_token=False)
        self.pos_embed.data.copy_(torch.from_numpy(pos_embed).float())
        torch.nn.init.normal_(self.mask_token, std=0.02)

    # ------------------------------------------------------------------ masking (swin_mae3d.py:1314-1382)
    def window_masking_3d(self, x, patch_size=(4, 4, 4), p_remove=0.50, mask_token=None, sampling_strategy="random"):
        """Stand-alone form with the reference signature: returns (masked tokens, mask (B,H,W,D,1) float)."""
        if sampling_strategy != "random" or len(set(patch_size)) != 1:
            raise ValueError("only the 'random' strategy with cubic blocks is reachable from the reference forward")
        B, H, W, D, C = x.shape
        m = torch.from_numpy(draw_block_mask((H, W, D), p_remove, patch_size[0])).to(x.device)
        mb = m.bool()[None, ..., None]
        fill = torch.zeros(C, device=x.device, dtype=x.dtype) if mask_token is None else mask_token.to(x.device)
        return torch.where(mb, fill.view(1, 1, 1, 1, C), x), mb.expand(B, H, W, D, 1).to(x.dtype)

    # ------------------------------------------------------------------ helpers with the reference names
    def patchify_3d(self, x, mask=None):
        """swin_mae3d.py:1384-1404 on a (N,4,R,R,R)-shaped tensor of any memory layout."""
        p = self.patch_size[0]
        assert x.shape[2] == x.shape[3] == x.shape[4] and x.shape[2] % p == 0
        n = x.shape[2] // p
        out = x.reshape(x.shape[0], 4, n, p, n, p, n, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(x.shape[0], n, n, n, p ** 3, 4)
        if mask is not None:
            m = mask.reshape(x.shape[0], 4, n, p, n, p, n, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(x.shape[0], n, n, n, p ** 3, 4)
            return out, m[..., 0].unsqueeze(-1).int()
        return out

    def transform(self, x):
        """swin_mae3d.py:1432-1448: list of (4,X,Y,Z) -> (B,4,R,R,R) batch + (B,3) extents (the pad mask, compactly)."""
        return NF.pad_grids(x, self.resolution)

    # ------------------------------------------------------------------ forward (swin_mae3d.py:1450-1505)
    def forward_encoder_ecoder(self, x):
        """x (B,4,R,R,R) -> (pred (B,4,R,R,R) view of channels-last memory, mask_patches (B,n,n,n,1) float)."""
        n = self.num_patches
        B = x.shape[0]
        m_np = draw_block_mask((n, n, n), self.masking_prob, 4)
        tok_mask = torch.from_numpy(m_np).to(x.device, non_blocking=True)
        self._tok_mask_u8 = tok_mask
        t = self.patch_partition(x, self.pos_embed.view(-1, self.embed_dim), tok_mask.view(-1), self.mask_token)
        feats = []
        for stage in self.stages:
            t = stage(t)
            feats.append(t)                       # already channels-last: the reference's permute().contiguous() is free
        d = self.decoder4.forward_cl(feats[3], feats[2])
        d = self.decoder3.forward_cl(d, feats[1])
        d = self.decoder2.forward_cl(d, feats[0])
        out = self.decoder1.forward_cl(d, out_block=self.out)      # decoder1 + the 1x1x1 output convolution (swin_mae3d.py:1494-1495)
        mask_patches = tok_mask.view(1, n, n, n, 1).expand(B, n, n, n, 1).to(x.dtype)
        return from_channels_last(out), mask_patches

    def forward_loss(self, x, pred, ext, mask_patches=None, is_eval=False):
        """swin_mae3d.py:1513-1563.  `ext` is the (B,3) extents tensor from transform(); `mask_patches` may be the
        float mask returned by forward_encoder_ecoder or None (then the mask of the last forward is used)."""
        if mask_patches is not None:
            tok = (mask_patches[0, ..., 0] != 0).to(torch.uint8).contiguous()
        else:
            tok = self._tok_mask_u8
        out3 = NF.MAELossFn.apply(to_channels_last(pred), x, ext, tok, self.patch_size[0])
        if not is_eval:
            return out3[0], out3[1], out3[2]
        target = self.patchify_3d(x)
        valid = target[..., 3].unsqueeze(-1) > 0.01
        return out3[0], out3[1], out3[2], self.patchify_3d(pred), valid, target

    def forward(self, x, is_eval=False):
        xb, ext = self.transform(x)
        return self.forward_padded(xb, ext, is_eval)

    def forward_padded(self, xb, ext, is_eval=False):
        """forward() on an already padded batch (B,4,R,R,R) + (B,3) extents, e.g. from functional.ingest_scenes()."""
        pred, mask_patches = self.forward_encoder_ecoder(xb)
        return self.forward_loss(xb, pred, ext, None, is_eval)


# NOTE: the reference driver imports `SwinTransformer_MAE3D_New as SwinTransformer_MAE3D` (run_swin_mae3d.py:22); the class that is
# actually NAMED SwinTransformer_MAE3D in the reference is the older model (swin_mae3d.py:417-1064): see swin_mae3d_legacy.py.

SWIN_CONFIGS = {  # run_swin_mae3d.py:378-399
    "swin_t": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]),
    "swin_s": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24]),
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32]),   # convention of SURVEY 8c
    "swin_l": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48]),
}


def build_model(backbone_type: str = "swin_s", resolution: int = 160, masking_prob: float = 0.75, **kw):
    cfg = SWIN_CONFIGS[backbone_type]
    return SwinTransformer_MAE3D_New(patch_size=[4, 4, 4], embed_dim=cfg["embed_dim"], depths=cfg["depths"],
                                     num_heads=cfg["num_heads"], window_size=[4, 4, 4], resolution=resolution,
                                     masking_prob=masking_prob, **kw)
