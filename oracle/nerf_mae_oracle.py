"""CPU oracle for the 3D Swin-MAE pretraining hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 *restatement* of the reference's algorithm
(zubair-irshad/NeRF-MAE @ 721b5ee).  It is the checker for the CUDA path and the
timed "reference arm" / ``cpu_baseline`` of ``bench.py``.  Only ``tests/``,
``__graft_entry__.smoke()`` and those two bench legs may import it; the product
package ``nerf-mae_b200`` never does.

Parity pinning: the reference ships no tests/golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the *live* reference
imported from ``/root/reference`` in the build container
(``tests/test_oracle_vs_reference.py``, skipped where the reference is absent)
and against the fixtures ``oracle/make_golden.py`` wrote from that live
reference into ``tests/golden/``.

Everything is written functionally over a ``state_dict`` that uses the
reference's parameter names, so reference checkpoints and the
product model's ``state_dict()`` can both be fed in unchanged.

Reference citations are ``file:line`` relative to ``/root/reference``; the three
files are abbreviated  S = nerf_mae/model/mae/swin_mae3d.py,
U = nerf_mae/model/mae/unetr_block.py, T = nerf_mae/model/mae/torch_utils.py.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SWIN_CONFIGS = {  # S:1603-1624 (and run_swin_mae3d.py:378-399)
    "swin_t": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]),
    "swin_s": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24]),
    # swin_b: the reference's own head list [3,6,12,24] does not divide C=128 and its
    # sincos table has 126 != 128 channels; convention of SURVEY 8(c).
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32]),
    "swin_l": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48]),
}


# ----------------------------------------------------------------------------- a3
def sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    """T:35-53.  [sin(p*w) | cos(p*w)], w_n = 10000^(-n/(dim/2)), fp64."""
    assert dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    out = pos.reshape(-1).astype(np.float64)[:, None] * omega[None, :]
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, n: int) -> Tensor:
    """T:5-32.  Returns (1,n,n,n,embed_dim) fp32.

    np.meshgrid defaults to 'xy' indexing (T:14), so for the table entry [i,j,k]
    the three channel groups encode (j, i, k) in that order.  Each group has
    embed_dim//3 channels (T:28-30); if 3*(embed_dim//3) < embed_dim the
    reference cannot be built at all (swin_b) - we zero-pad the tail.
    """
    per = embed_dim // 3
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    emb = np.concatenate(
        [sincos_1d(per, j.astype(np.float32)), sincos_1d(per, i.astype(np.float32)),
         sincos_1d(per, k.astype(np.float32))], axis=1)
    if emb.shape[1] < embed_dim:
        emb = np.concatenate([emb, np.zeros((emb.shape[0], embed_dim - emb.shape[1]))], axis=1)
    return torch.from_numpy(emb.reshape(1, n, n, n, embed_dim)).float()


# ----------------------------------------------------------------------------- a1
def pad_grids(grids: Sequence[Tensor], R: int) -> Tuple[Tensor, Tensor]:
    """T:56-90 + S:1432-1448,1572-1574.  list of (4,X,Y,Z) -> (B,4,R,R,R) zero padded
    at the high end of every axis, and the extents (B,3) int64 that replace the
    reference's dense 0/1 pad mask (mask[b,:,x,y,z] = x<X and y<Y and z<Z)."""
    out = torch.zeros(len(grids), 4, R, R, R, dtype=grids[0].dtype)
    ext = torch.zeros(len(grids), 3, dtype=torch.int64)
    for b, g in enumerate(grids):
        _, X, Y, Z = g.shape
        out[b, :, :X, :Y, :Z] = g
        ext[b] = torch.tensor([X, Y, Z])
    return out, ext


# ----------------------------------------------------------------------------- a4
def draw_block_mask(n_tok: Sequence[int], p_remove: float, block: int = 4, rng=random) -> Tensor:
    """S:1364-1373.  One ``rng.random() < p`` per block^3 block of *tokens*, h-major,
    d-minor; blocks start at 0,block,.. while start <= n-block.  Returns bool
    (H,W,D) at token resolution (True = masked).  Shared by every grid of the batch."""
    H, W, D = n_tok
    m = torch.zeros(H, W, D, dtype=torch.bool)
    for h in range(0, H - block + 1, block):
        for w in range(0, W - block + 1, block):
            for d in range(0, D - block + 1, block):
                if rng.random() < p_remove:
                    m[h:h + block, w:w + block, d:d + block] = True
    return m


# ----------------------------------------------------------------------------- norms / acts
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def instance_norm_cl(x: Tensor, eps: float = 1e-5) -> Tensor:
    """U:77 nn.InstanceNorm3d: no affine, biased variance, eps 1e-5; x is (B,C,H,W,D)."""
    mu = x.mean(dim=(2, 3, 4), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3, 4), keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


def leaky_relu(x: Tensor, slope: float = 0.01) -> Tensor:  # U:83
    return torch.where(x >= 0, x, x * slope)


# ----------------------------------------------------------------------------- a2
def patch_embed(x: Tensor, sd: Dict[str, Tensor], p: int = 4) -> Tensor:
    """S:1120-1129.  Conv3d(k=s=p) restated as a per-patch GEMM, then LayerNorm.
    (B,4,R,R,R) -> (B,R/p,R/p,R/p,C)."""
    B, Ci, R, _, _ = x.shape
    n = R // p
    w = sd["patch_partition.0.weight"]          # (C,Ci,p,p,p)
    C = w.shape[0]
    cols = x.reshape(B, Ci, n, p, n, p, n, p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, n, n, n, Ci * p ** 3)
    y = cols @ w.reshape(C, -1).t() + sd["patch_partition.0.bias"]
    return layer_norm(y, sd["patch_partition.2.weight"], sd["patch_partition.2.bias"])


# ----------------------------------------------------------------------------- a5/a6
def relative_position_index(ws: int = 4) -> Tensor:
    """S:257-280.  index[q,k] = ((dh+ws-1)*(2ws-1) + (dw+ws-1))*(2ws-1) + (dd+ws-1),
    delta = coord(q) - coord(k), tokens of a window flattened h-major."""
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    d = c[:, :, None] - c[:, None, :] + (ws - 1)
    return ((d[0] * (2 * ws - 1) + d[1]) * (2 * ws - 1) + d[2]).flatten()


def window_token_map(H: int, W: int, D: int, ws: int, shift: int):
    """Index-remap formulation of pad -> roll -> window-partition (S:60-101).

    Returns (src, region): for every (window, slot) the flat index of the source
    token in the un-padded (H,W,D) grid, or -1 when the slot is a padding token,
    and the shift-mask region id (S:126-158) of the slot.  ``shift`` is already
    zeroed where window >= padded size (S:68-75; cubic grids -> all-or-nothing)."""
    P = [((n + ws - 1) // ws) * ws for n in (H, W, D)]
    sh = [0 if ws >= Pn else shift for Pn in P]
    if sum(sh) == 0:
        sh = [0, 0, 0]
    ax = []
    for n, Pn, s in zip((H, W, D), P, sh):
        r = torch.arange(Pn)                 # coordinate in the rolled frame
        srcc = (r + s) % Pn                  # torch.roll(x, -s): rolled[r] = x[(r+s) % P]
        band = torch.zeros(Pn, dtype=torch.long)
        if s > 0:
            band[Pn - ws:Pn - s] = 1
            band[Pn - s:] = 2
        ax.append((srcc, band, n))
    (sh_, bh, _), (sw_, bw, _), (sd_, bd, _) = ax
    valid = (sh_[:, None, None] < H) & (sw_[None, :, None] < W) & (sd_[None, None, :] < D)
    flat = (sh_[:, None, None] * W + sw_[None, :, None]) * D + sd_[None, None, :]
    src = torch.where(valid, flat, torch.full_like(flat, -1))
    region = (bh[:, None, None] * 3 + bw[None, :, None]) * 3 + bd[None, None, :]

    def part(t):
        t = t.reshape(P[0] // ws, ws, P[1] // ws, ws, P[2] // ws, ws)
        return t.permute(0, 2, 4, 1, 3, 5).reshape(-1, ws ** 3)
    return part(src), part(region), sum(sh) > 0


def window_attention(x: Tensor, qkv_w: Tensor, qkv_b: Optional[Tensor], proj_w: Tensor, proj_b: Optional[Tensor],
                     bias_table: Tensor, num_heads: int, ws: int = 4, shift: int = 0) -> Tensor:
    """S:27-197 restated per token: QKV is a per-token linear so it is evaluated in
    token order; padding tokens (zeros after LN, S:62-65) therefore carry q/k/v
    equal to the qkv bias and are NOT masked out of the softmax."""
    B, H, W, D, C = x.shape
    hd = C // num_heads
    N = ws ** 3
    src, region, shifted = window_token_map(H, W, D, ws, shift)
    nW = src.shape[0]
    qkv = F.linear(x.reshape(B, H * W * D, C), qkv_w, qkv_b)                 # S:108
    pad_row = qkv_b if qkv_b is not None else torch.zeros(3 * C)
    qkv = torch.cat([qkv, pad_row.expand(B, 1, 3 * C)], dim=1)              # slot -1 -> bias row
    g = qkv[:, src.reshape(-1)].reshape(B, nW, N, 3, num_heads, hd)
    q, k, v = (g[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))      # (B,nW,nH,N,hd)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)                           # S:119-120
    rpb = bias_table[relative_position_index(ws)].reshape(N, N, num_heads).permute(2, 0, 1)  # S:200-211
    attn = attn + rpb
    if shifted:                                                             # S:124-167
        diff = region[:, :, None] != region[:, None, :]
        attn = attn + (diff.float() * -100.0)[None, :, None]
    attn = torch.softmax(attn, dim=-1)                                      # S:169
    o = (attn @ v).permute(0, 1, 3, 2, 4).reshape(B, nW * N, C)             # S:172
    o = F.linear(o, proj_w, proj_b)                                         # S:173
    out = torch.zeros(B, H * W * D, C)
    keep = src.reshape(-1) >= 0                                             # crop of padded queries, S:196
    out[:, src.reshape(-1)[keep]] = o[:, keep]
    return out.reshape(B, H, W, D, C)


# ----------------------------------------------------------------------------- a7
def swin_block(x: Tensor, sd: Dict[str, Tensor], pre: str, num_heads: int, shift: int,
               sd_scale: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """S:366-369.  ``sd_scale`` = per-sample stochastic-depth multipliers for the two
    branches ((B,) each, 0 or 1/(1-p)); None = eval."""
    a = window_attention(layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]),
                         sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"],
                         sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"],
                         sd[pre + "attn.relative_position_bias_table"], num_heads, 4, shift)
    if sd_scale is not None:
        a = a * sd_scale[0].view(-1, 1, 1, 1, 1)
    x = x + a
    h = layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    h = gelu_erf(F.linear(h, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"]))
    h = F.linear(h, sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"])
    if sd_scale is not None:
        h = h * sd_scale[1].view(-1, 1, 1, 1, 1)
    return x + h


# ----------------------------------------------------------------------------- a8
def patch_merge(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """S:390-414.  Channel block index of the 2x2x2 neighbour (dh,dw,dd) is dh+2dw+4dd."""
    B, H, W, D, C = x.shape
    x = F.pad(x, (0, 0, 0, D % 2, 0, W % 2, 0, H % 2))
    H2, W2, D2 = x.shape[1] // 2, x.shape[2] // 2, x.shape[3] // 2
    x = x.reshape(B, H2, 2, W2, 2, D2, 2, C).permute(0, 1, 3, 5, 6, 4, 2, 7).reshape(B, H2, W2, D2, 8 * C)
    x = layer_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    return F.linear(x, sd[pre + "reduction.weight"])


# ----------------------------------------------------------------------------- a10-a12
def conv_transpose_k_eq_s(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """U:151-158 with kernel == stride, padding 0: every input voxel owns a disjoint
    k^3 block of outputs, so it is a per-voxel GEMM + depth-to-space.  x (B,Ci,H,W,D),
    w (Ci,Co,k,k,k)."""
    B, Ci, H, W, D = x.shape
    Co, k = w.shape[1], w.shape[2]
    y = x.permute(0, 2, 3, 4, 1).reshape(-1, Ci) @ w.reshape(Ci, Co * k ** 3)
    y = y.reshape(B, H, W, D, Co, k, k, k).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(B, Co, H * k, W * k, D * k)
    return y + b.view(1, -1, 1, 1, 1)


def res_block(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """U:57-71 (UnetResBlock, instance norm, LeakyReLU 0.01)."""
    out = F.conv3d(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    out = leaky_relu(instance_norm_cl(out))
    out = F.conv3d(out, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    out = instance_norm_cl(out)
    res = x
    if pre + "conv3.weight" in sd:
        res = instance_norm_cl(F.conv3d(x, sd[pre + "conv3.weight"], sd[pre + "conv3.bias"]))
    return leaky_relu(out + res)


def up_block(x: Tensor, skip: Optional[Tensor], sd: Dict[str, Tensor], pre: str) -> Tensor:
    """U:193-200."""
    out = conv_transpose_k_eq_s(x, sd[pre + "transp_conv.weight"], sd[pre + "transp_conv.bias"])
    if skip is not None:
        out = torch.cat((out, skip), dim=1)
    return res_block(out, sd, pre + "conv_block.")


# ----------------------------------------------------------------------------- a13/a14
def patchify(x: Tensor, p: int = 4) -> Tensor:
    """S:1384-1394.  (N,4,R,R,R) -> (N,h,w,l,p^3,4)."""
    N, C, R = x.shape[0], x.shape[1], x.shape[2]
    n = R // p
    return x.reshape(N, C, n, p, n, p, n, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(N, n, n, n, p ** 3, C)


def mae_loss(x: Tensor, pred: Tensor, ext: Tensor, tok_mask: Tensor):
    """S:1513-1563.  x,pred (B,4,R,R,R); ext (B,3) un-padded extents; tok_mask bool
    (h,w,l) token-level block mask shared over the batch.

    loss_rgb   = sum_{vox,3ch} (p-t)^2 [t_a > 0.01] / #[t_a > 0.01]         (S:1529-1535)
    loss_alpha = sum_vox (sigmoid(p_a)-t_a)^2 in_extent*masked / sum(in_extent*masked)
    """
    B, _, R, _, _ = x.shape
    valid = x[:, 3] > 0.01
    loss_rgb = (((pred[:, :3] - x[:, :3]) ** 2) * valid[:, None]).sum() / valid.sum()
    ar = torch.arange(R)
    inb = ((ar[None, :, None, None] < ext[:, 0, None, None, None]) &
           (ar[None, None, :, None] < ext[:, 1, None, None, None]) &
           (ar[None, None, None, :] < ext[:, 2, None, None, None]))
    p = R // tok_mask.shape[0]
    vm = tok_mask.repeat_interleave(p, 0).repeat_interleave(p, 1).repeat_interleave(p, 2)
    rem = (inb & vm[None]).float()
    loss_alpha = (((torch.sigmoid(pred[:, 3]) - x[:, 3]) ** 2) * rem).sum() / rem.sum()
    return loss_rgb + loss_alpha, loss_rgb, loss_alpha, valid


# ----------------------------------------------------------------------------- model
def encoder(tokens: Tensor, sd: Dict[str, Tensor], depths: Sequence[int], num_heads: Sequence[int],
            sd_scales=None) -> List[Tensor]:
    """S:1466-1470.  tokens (B,h,w,l,C) -> the 4 stage outputs (NDHWC)."""
    feats, blk = [], 0
    x = tokens
    for s, depth in enumerate(depths):
        off = 0
        if s > 0:
            x = patch_merge(x, sd, f"stages.{s}.0.")
            off = 1
        for i in range(depth):
            x = swin_block(x, sd, f"stages.{s}.{i + off}.", num_heads[s], 0 if i % 2 == 0 else 2,
                           None if sd_scales is None else sd_scales[blk])
            blk += 1
        feats.append(x)
    return feats


def forward(sd: Dict[str, Tensor], grids: Sequence[Tensor], depths: Sequence[int], num_heads: Sequence[int],
            resolution: int, masking_prob: float, is_eval: bool = False, rng=random,
            tok_mask: Optional[Tensor] = None, sd_scales=None, return_internals: bool = False):
    """S:1571-1599 (forward) + S:1450-1505 (forward_encoder_ecoder)."""
    x, ext = pad_grids(grids, resolution)
    t = patch_embed(x, sd) + sd["pos_embed"]                                  # S:1455-1459
    if tok_mask is None:
        tok_mask = draw_block_mask(t.shape[1:4], masking_prob, 4, rng)        # S:1461
    t = torch.where(tok_mask[None, ..., None], sd["mask_token"].view(1, 1, 1, 1, -1), t)  # S:1375-1380
    feats = encoder(t, sd, depths, num_heads, sd_scales)
    f = [v.permute(0, 4, 1, 2, 3) for v in feats]                             # S:1470
    d3 = up_block(f[3], f[2], sd, "decoder4.")
    d2 = up_block(d3, f[1], sd, "decoder3.")
    d1 = up_block(d2, f[0], sd, "decoder2.")
    d0 = up_block(d1, None, sd, "decoder1.")
    pred = F.conv3d(d0, sd["out.conv.weight"], sd["out.conv.bias"])           # S:1495
    loss, loss_rgb, loss_alpha, valid = mae_loss(x, pred, ext, tok_mask)
    if return_internals:
        return dict(loss=loss, loss_rgb=loss_rgb, loss_alpha=loss_alpha, pred=pred, tokens=t, feats=feats,
                    dec=[d3, d2, d1, d0], tok_mask=tok_mask, x=x, ext=ext)
    if is_eval:
        return loss, loss_rgb, loss_alpha, patchify(pred), patchify(valid[:, None].expand(-1, 4, -1, -1, -1))[..., :1], patchify(x)
    return loss, loss_rgb, loss_alpha


def init_state_dict(name: str, resolution: int, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init parameters with the reference's shapes and init *distributions*
    (S:1272-1276, 255, 1312); used where bit-identical reference init is not needed
    (bench reference arm)."""
    cfg = SWIN_CONFIGS[name]
    C0, depths, heads = cfg["embed_dim"], cfg["depths"], cfg["num_heads"]
    g = torch.Generator().manual_seed(seed)

    def tn(*shape, std=0.02):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=std, generator=g)

    def conv(co, ci, k):
        bound = 1.0 / math.sqrt(ci * k ** 3)
        return (torch.empty(co, ci, k, k, k).uniform_(-bound, bound, generator=g),
                torch.empty(co).uniform_(-bound, bound, generator=g))
    sd: Dict[str, Tensor] = {}
    n = resolution // 4
    sd["pos_embed"] = sincos_pos_embed_3d(C0, n)
    sd["mask_token"] = torch.empty(C0).normal_(0, 0.02, generator=g)
    sd["patch_partition.0.weight"], sd["patch_partition.0.bias"] = conv(C0, 4, 4)
    sd["patch_partition.2.weight"], sd["patch_partition.2.bias"] = torch.ones(C0), torch.zeros(C0)
    for s, depth in enumerate(depths):
        C = C0 * 2 ** s
        off = 0
        if s > 0:
            Ci = C // 2
            sd[f"stages.{s}.0.reduction.weight"] = tn(C, 8 * Ci)
            sd[f"stages.{s}.0.norm.weight"], sd[f"stages.{s}.0.norm.bias"] = torch.ones(8 * Ci), torch.zeros(8 * Ci)
            off = 1
        for i in range(depth):
            p = f"stages.{s}.{i + off}."
            for nm in ("norm1", "norm2"):
                sd[p + nm + ".weight"], sd[p + nm + ".bias"] = torch.ones(C), torch.zeros(C)
            sd[p + "attn.relative_position_bias_table"] = tn(343, heads[s])
            sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = tn(3 * C, C), torch.zeros(3 * C)
            sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = tn(C, C), torch.zeros(C)
            sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"] = tn(4 * C, C), torch.zeros(4 * C)
            sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"] = tn(C, 4 * C), torch.zeros(C)
    for name_, ci, co, k, skip in (("decoder4", 8 * C0, 4 * C0, 2, True), ("decoder3", 4 * C0, 2 * C0, 2, True),
                                   ("decoder2", 2 * C0, C0, 2, True), ("decoder1", C0, C0 // 2, 4, False)):
        bound = 1.0 / math.sqrt(co * k ** 3)
        sd[name_ + ".transp_conv.weight"] = torch.empty(ci, co, k, k, k).uniform_(-bound, bound, generator=g)
        sd[name_ + ".transp_conv.bias"] = torch.empty(co).uniform_(-bound, bound, generator=g)
        cin = 2 * co if skip else co
        sd[name_ + ".conv_block.conv1.weight"], sd[name_ + ".conv_block.conv1.bias"] = conv(co, cin, 3)
        sd[name_ + ".conv_block.conv2.weight"], sd[name_ + ".conv_block.conv2.bias"] = conv(co, co, 3)
        if skip:
            sd[name_ + ".conv_block.conv3.weight"], sd[name_ + ".conv_block.conv3.bias"] = conv(co, cin, 1)
    sd["out.conv.weight"], sd["out.conv.bias"] = conv(4, C0 // 2, 1)
    return sd


def train_step(sd: Dict[str, Tensor], opt_state: Dict, grids, depths, num_heads, resolution, masking_prob,
               lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, wd: float = 1e-3,
               clip: float = 0.1, rng=random, tok_mask=None, sd_scales=None):
    """One optimiser step as run_swin_mae3d.py:650-669 does it: fwd, bwd, global-L2
    clip (torch.nn.utils.clip_grad_norm_: scale by clip/(norm+1e-6) if < 1), AdamW
    (decoupled weight decay, bias-corrected).  ``sd`` tensors that require grad are
    updated in place; returns (loss, loss_rgb, loss_alpha, grad_norm)."""
    params = {k: v for k, v in sd.items() if v.requires_grad}
    for v in params.values():
        v.grad = None
    loss, lrgb, lalpha = forward(sd, grids, depths, num_heads, resolution, masking_prob, False, rng, tok_mask, sd_scales)
    loss.backward()
    gn = torch.sqrt(sum((v.grad.double() ** 2).sum() for v in params.values())).float()
    coef = torch.clamp(clip / (gn + 1e-6), max=1.0)
    opt_state["step"] = opt_state.get("step", 0) + 1
    t = opt_state["step"]
    with torch.no_grad():
        for k, v in params.items():
            g = v.grad * coef
            m = opt_state.setdefault("m." + k, torch.zeros_like(v))
            s = opt_state.setdefault("v." + k, torch.zeros_like(v))
            v.mul_(1 - lr * wd)
            m.mul_(beta1).add_(g, alpha=1 - beta1)
            s.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (s.sqrt() / math.sqrt(1 - beta2 ** t)).add_(eps)
            v.addcdiv_(m, denom, value=-lr / (1 - beta1 ** t))
    return loss.detach(), lrgb.detach(), lalpha.detach(), gn


# ----------------------------------------------------------------------------- a16 (BASELINE config 5)
def fpn_forward(sd: Dict[str, Tensor], feats: Sequence[Tensor], pre: str = "") -> List[Tensor]:
    """nerf_rpn/model/fpn.py:134-185 on its default path (no extra convs, num_outs == number of inputs).
    feats: (B,C_i,s_i,s_i,s_i) tensors; returns the (B,256,s_i,s_i,s_i) pyramid."""
    n = len(feats)
    lat = [F.conv3d(feats[i], sd[f"{pre}lateral_convs.{i}.weight"], sd[f"{pre}lateral_convs.{i}.bias"]) for i in range(n)]  # :139-142
    for i in range(n - 1, 0, -1):                                                                                      # :146-158
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    return [F.conv3d(lat[i], sd[f"{pre}fpn_convs.{i}.weight"], sd[f"{pre}fpn_convs.{i}.bias"], padding=1) for i in range(n)]  # :162-164


def encoder_features(sd: Dict[str, Tensor], x: Tensor, depths: Sequence[int], num_heads: Sequence[int]) -> List[Tensor]:
    """nerf_rpn/model/feature_extractor.py:1171-1184: patch_partition + pos_embed (no masking) -> stage outputs as
    (B,C,H,W,D) tensors.  x: (B,4,R,R,R)."""
    t = patch_embed(x, sd) + sd["pos_embed"]
    return [v.permute(0, 4, 1, 2, 3) for v in encoder(t, sd, depths, num_heads, None)]
