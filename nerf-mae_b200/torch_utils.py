"""Host-side helpers mirroring nerf_mae/model/mae/torch_utils.py of the reference (same names, same results)."""
from __future__ import annotations

import numpy as np
import torch

from . import functional as NF


def get_1d_sincos_pos_embed_from_grid(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """torch_utils.py:35-53: [sin(p w) | cos(p w)], w_n = 10000^(-n/(D/2)), evaluated in float64."""
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1).astype(np.float64), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_3d_sincos_pos_embed(embed_dim: int, grid_size: int, cls_token: bool = False) -> np.ndarray:
    """torch_utils.py:5-32.  (1,n,n,n,embed_dim).  np.meshgrid's default 'xy' indexing of the reference makes the
    first channel group encode the SECOND token axis: table[i,j,k] = [f(j) | f(i) | f(k)].
    embed_dim not divisible by 3 (swin_b, C=128) cannot be built by the reference at all; the tail channels
    are zero here (convention of SURVEY 8c)."""
    n = grid_size
    per = embed_dim // 3
    i, j, k = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32),
                          indexing="ij")
    emb = np.concatenate([get_1d_sincos_pos_embed_from_grid(per, j), get_1d_sincos_pos_embed_from_grid(per, i),
                          get_1d_sincos_pos_embed_from_grid(per, k)], axis=1)
    if emb.shape[1] < embed_dim:
        emb = np.concatenate([emb, np.zeros((emb.shape[0], embed_dim - emb.shape[1]))], axis=1)
    emb = emb.reshape(1, n, n, n, embed_dim)
    if cls_token:
        raise NotImplementedError("cls_token is never used by the MAE path")
    return emb


def pad_tensor(tensor: torch.Tensor, target_shape, pad_value: float = 0):
    """torch_utils.py:56-90 for one (4,X,Y,Z) grid: returns the (1,4,R,R,R) padded grid and, instead of the
    reference's dense CPU 0/1 mask, the (1,3) int32 un-padded extents."""
    if pad_value != 0:
        raise NotImplementedError("the MAE path only pads with zeros")
    R = int(target_shape[0])
    return NF.pad_grids([tensor], R)
