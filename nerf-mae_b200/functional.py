"""Autograd operators of the B200 3D Swin-MAE path.  Every forward/backward is a sequence of
C-ABI calls into libnmae.so (include/nmae.h); torch only provides device memory and streams.

Token tensors are channels-last (B,H,W,D,C) as in the reference encoder; decoder volumes are kept
channels-last (B,X,Y,Z,C) in memory and exposed as (B,C,X,Y,Z) views at the module boundary.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

import weakref

import numpy as np

from ._lib import call, conv3_image_bytes, conv3h_image_bytes, linear_blob_layout, num_windows, workspace_bytes


# Operand precision of the decoder's 3x3x3 convolutions on the tensor cores:
#   "bf16x3": each fp32 operand split into bf16 hi + lo, three products per term (fp32-class, ~2^-17 relative);
#   "fp16"  : one pass over fp16 operands (11-bit significands like the TF32 the reference's cuDNN convolutions use on a GPU
#             by default), fp32 accumulation; output-gradient images carry a per-tensor power-of-two scale.  The default: it is
#             the numerical class of the reference's own GPU training path and meets the 1e-3 known-answer tests at every
#             BASELINE size (tests/test_gpu_sized.py); "bf16x3" is the strict mode.
CONV_PRECISIONS = ("bf16x3", "fp16")
DEFAULT_CONV_PRECISION = "fp16"
_conv_precision = DEFAULT_CONV_PRECISION


def set_conv_precision(mode: Optional[str]) -> str:
    """Select the operand precision of the 3x3x3 convolutions (None restores the default); returns the previous mode."""
    global _conv_precision
    prev = _conv_precision
    mode = DEFAULT_CONV_PRECISION if mode is None else mode
    if mode not in CONV_PRECISIONS:
        raise ValueError(f"conv precision must be one of {CONV_PRECISIONS}, got {mode!r}")
    _conv_precision = mode
    return prev


def get_conv_precision() -> str:
    return _conv_precision


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"nerf-mae_b200 computes in fp32 (the reference is fp32-only); got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _empty(ref: torch.Tensor, *shape, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device=ref.device)


# --------------------------------------------------------------------------------------------- weight blobs
class _WeightBlobs:
    """Tensor-core weight blobs of the linear layers, kept across calls (include/nmae.h: nmae_linear_prep_batch).

    The GEMM kernels read weights as bf16 hi/lo "blobs" in their shared-memory layout.  Re-laying a weight costs one small kernel;
    a Swin-S step has 212 such launches.  This registry keeps one persistent blob per (weight, GEMM shape, direction), and
    `refresh()` - called by FusedAdamWClip.step() after the parameters changed - rebuilds ALL of them with one launch.  A blob is
    used without re-laying only while (a) no refresh-less weight update can have happened: the entry's epoch is the registry's and
    the tensor's torch version counter is unchanged, and (b) the entry still refers to the very same tensor object.  Code that
    changes weights behind torch's back (`.data` tricks, raw pointers) must call invalidate()."""

    def __init__(self):
        self.entries = {}
        self.epoch = 0

    def invalidate(self):
        self.epoch += 1

    def get(self, w: torch.Tensor, M: int, N: int, K: int, transposed: bool):
        """Blob buffer for out[M,N] = A[M,K] W^T with W = w (transposed=False: forward) or w^T (input gradient), and the flag
        bits for the C call (256 when the blob is current)."""
        key = (w.data_ptr(), M, N, K, transposed)
        e = self.entries.get(key)
        if e is None or e["w"]() is not w:
            nt, kg = linear_blob_layout(M, N, K, w.device.index if w.device.index is not None else torch.cuda.current_device())
            if nt == 0:                                   # CUDA-core path: plain scratch, nothing to keep
                return torch.empty(w.numel(), dtype=torch.float32, device=w.device), 0
            s_n, s_k = (1, N) if transposed else (K, 1)    # W(n,k) = w[n*s_n + k*s_k]; transposed: w is (K_gemm, N_gemm) row-major
            e = dict(w=weakref.ref(w), blob=torch.empty(w.numel(), dtype=torch.float32, device=w.device), nt=nt, kg=kg, N=N, K=K,
                     s_n=s_n, s_k=s_k, epoch=-1, version=-1)
            self.entries[key] = e
        fresh = e["epoch"] == self.epoch and e["version"] == w._version
        e["epoch"], e["version"] = self.epoch, w._version   # the call about to be made (re)builds the blob when it is not fresh
        return e["blob"], (256 if fresh else 0)

    def is_current(self, w: torch.Tensor, M: int, N: int, K: int, transposed: bool) -> bool:
        """Whether the next get() for this use would skip the re-lay (no side effects: for tests and diagnostics)."""
        e = self.entries.get((w.data_ptr(), M, N, K, transposed))
        return e is not None and e["w"]() is w and e["epoch"] == self.epoch and e["version"] == w._version

    def refresh(self):
        """Weights were updated in place by the optimizer kernel: rebuild every live blob with one launch per device."""
        self.epoch += 1
        by_dev = {}
        for key, e in list(self.entries.items()):
            w = e["w"]()
            if w is None or w.data_ptr() != key[0]:
                del self.entries[key]
                continue
            by_dev.setdefault(w.device, []).append((w, e))
        for dev, items in by_dev.items():
            tbl = np.asarray([[w.data_ptr(), e["blob"].data_ptr(), e["N"], e["K"], e["nt"], e["s_n"], e["s_k"], e["kg"]]
                              for w, e in items], dtype=np.int64)
            dtab = torch.from_numpy(tbl).pin_memory().to(dev, non_blocking=True)
            call("nmae_linear_prep_batch", dtab, len(items), int(max(e["N"] * e["K"] for _, e in items)), device=dev)
            for w, e in items:
                e["epoch"], e["version"] = self.epoch, w._version


weight_blobs = _WeightBlobs()


# --------------------------------------------------------------------------------------------- patch embed
class PatchEmbedFn(Function):
    """patch_partition (Conv3d k=s=p + LayerNorm) [+ pos_embed + mask-token replacement]; swin_mae3d.py:1120-1129,1455-1463."""

    @staticmethod
    def forward(ctx, x, w, b, ln_w, ln_b, eps, pos, mask_u8, mask_token, p):
        x = _f32c(x)
        B, Ci, R = x.shape[0], x.shape[1], x.shape[2]
        if Ci != 4 or x.shape[3] != R or x.shape[4] != R:
            raise ValueError(f"patch embed expects (B,4,R,R,R) grids, got {tuple(x.shape)}")
        C = w.shape[0]
        n = R // p
        T = n ** 3
        w, b, ln_w, ln_b = _f32c(w), _f32c(b), _f32c(ln_w), _f32c(ln_b)
        conv = _empty(x, B * T, C)
        mean, rstd = _empty(x, B * T), _empty(x, B * T)
        tokens = _empty(x, B, n, n, n, C)
        if pos is not None:
            pos = _f32c(pos)
        if mask_u8 is not None:
            mask_u8 = mask_u8.contiguous()
            mask_token = _f32c(mask_token)
        call("nmae_patch_embed_fwd", x, w, b, ln_w, ln_b, pos, mask_u8, mask_token if mask_u8 is not None else None,
             B, R, p, C, float(eps), conv, mean, rstd, tokens, _empty(x, w.numel()), device=x.device)
        ctx.save_for_backward(x, w, ln_w, conv, mean, rstd, mask_u8)
        ctx.dims = (B, R, p, C)
        ctx.has_mask = mask_u8 is not None
        return tokens

    @staticmethod
    @once_differentiable
    def backward(ctx, dtok):
        x, w, ln_w, conv, mean, rstd, mask_u8 = ctx.saved_tensors
        B, R, p, C = ctx.dims
        dtok = _f32c(dtok)
        T = (R // p) ** 3
        ws = _empty(x, B * T, C)
        dw = torch.empty_like(w)
        db, dlw, dlb, dmt = (_empty(x, C) for _ in range(4))
        call("nmae_patch_embed_bwd", dtok, x, w, ln_w, conv, mean, rstd, mask_u8, B, R, p, C, ws, dw, db, dlw, dlb, dmt,
             device=x.device)
        return None, dw, db, dlw, dlb, None, None, None, (dmt if ctx.has_mask else None), None


# --------------------------------------------------------------------------------------------- LayerNorm / Linear
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        x, w, b = _f32c(x), _f32c(w), _f32c(b)
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        mean, rstd = _empty(x, rows), _empty(x, rows)
        call("nmae_layernorm_fwd", x, w, b, rows, C, float(eps), y, mean, rstd, device=x.device)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = _f32c(dy)
        C = x.shape[-1]
        rows = x.numel() // C
        dx = torch.empty_like(x)
        dw, db = torch.empty_like(w), torch.empty_like(w)
        call("nmae_layernorm_bwd", dy, x, w, mean, rstd, rows, C, None, dx, dw, db, device=x.device)
        return dx, dw, db, None


class LinearFn(Function):
    """F.linear over the last dim."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _f32c(x), _f32c(w)
        K, N = x.shape[-1], w.shape[0]
        M = x.numel() // K
        out = _empty(x, *x.shape[:-1], N)
        ws, rdy = weight_blobs.get(w, M, N, K, False)
        call("nmae_linear_fwd", x, w, None if b is None else _f32c(b), M, N, K, rdy, None, None, None, 1, out, ws, device=x.device)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32c(dy)
        K, N = x.shape[-1], w.shape[0]
        M = x.numel() // K
        dx = torch.empty_like(x)
        ws, rdy = weight_blobs.get(w, M, K, N, True)
        call("nmae_linear_bwd_input", dy, w, M, N, K, rdy, None, dx, ws, device=x.device)
        dw = torch.empty_like(w)
        db = _empty(x, N) if ctx.has_bias else None
        call("nmae_linear_bwd_weight", dy, x, M, N, K, dw, db, device=x.device)
        return dx, dw, db


# --------------------------------------------------------------------------------------------- W-MSA
class WindowAttentionFn(Function):
    """[LayerNorm ->] qkv -> shifted-window attention (window 4^3, rel-pos bias, shift mask) -> proj
    [-> x + row_scale * .];  swin_mae3d.py:27-197 (+ :366 when fused with norm1 / residual)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, qkv_w, qkv_b, proj_w, proj_b, table, num_heads, shift, residual, row_scale):
        x = _f32c(x)
        B, H, W, D, C = x.shape
        T = H * W * D
        M = B * T
        dev = x.device
        qkv_w, proj_w, table = _f32c(qkv_w), _f32c(proj_w), _f32c(table)
        if ln_w is not None:
            ln_w, ln_b = _f32c(ln_w), _f32c(ln_b)
            h = torch.empty_like(x)
            mean, rstd = _empty(x, M), _empty(x, M)
            call("nmae_layernorm_fwd", x, ln_w, ln_b, M, C, float(eps), h, mean, rstd, device=dev)
        else:
            h, mean, rstd = x, None, None
        qkv = _empty(x, M + 1, 3 * C)
        ws, rdy = weight_blobs.get(qkv_w, M, 3 * C, C, False)
        call("nmae_linear_fwd", h, qkv_w, qkv_b, M, 3 * C, C, rdy, None, None, None, 1, qkv, ws, device=dev)
        if qkv_b is not None:            # padding tokens are zeros before the projection -> their q/k/v are the bias
            qkv[M].copy_(qkv_b)
        else:
            qkv[M].zero_()
        nW = num_windows(H, W, D)
        attn = _empty(x, M, C)
        lse = _empty(x, B * nW * num_heads * 64)
        call("nmae_window_attention_fwd", qkv, table, B, H, W, D, C, num_heads, shift, attn, lse, device=dev)
        out = torch.empty_like(x)
        if row_scale is not None:
            row_scale = _f32c(row_scale)
        ws, rdy = weight_blobs.get(proj_w, M, C, C, False)
        call("nmae_linear_fwd", attn, proj_w, proj_b, M, C, C, (2 if residual else 0) | rdy, None, x if residual else None,
             row_scale if residual else None, T, out, ws, device=dev)
        ctx.save_for_backward(x, ln_w, mean, rstd, h if ln_w is not None else None, qkv_w, proj_w, table, qkv, attn, lse, row_scale)
        ctx.cfg = (num_heads, shift, residual, qkv_b is not None, proj_b is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, ln_w, mean, rstd, h, qkv_w, proj_w, table, qkv, attn, lse, row_scale = ctx.saved_tensors
        num_heads, shift, residual, has_qb, has_pb = ctx.cfg
        dout = _f32c(dout)
        B, H, W, D, C = x.shape
        T = H * W * D
        M = B * T
        dev = x.device
        if h is None:
            h = x
        dproj = dout
        if residual and row_scale is not None:
            dproj = torch.empty_like(dout)
            call("nmae_scale_rows", dproj, dout, row_scale, T, M, C, device=dev)
        dattn = _empty(x, M, C)
        ws, rdy = weight_blobs.get(proj_w, M, C, C, True)
        call("nmae_linear_bwd_input", dproj, proj_w, M, C, C, rdy, None, dattn, ws, device=dev)
        dpw = torch.empty_like(proj_w)
        dpb = _empty(x, C) if has_pb else None
        call("nmae_linear_bwd_weight", dproj, attn, M, C, C, dpw, dpb, device=dev)
        dqkv = _empty(x, M + 1, 3 * C)
        dtable = torch.empty_like(table)
        call("nmae_window_attention_bwd", dattn, qkv, table, attn, lse, B, H, W, D, C, num_heads, shift, dqkv, dtable, device=dev)
        dqw = torch.empty_like(qkv_w)
        call("nmae_linear_bwd_weight", dqkv, h, M, 3 * C, C, dqw, None, device=dev)
        dqb = None
        if has_qb:
            dqb = _empty(x, 3 * C)
            call("nmae_colsum", dqkv, M + 1, 3 * C, 3 * C, dqb, device=dev)   # row M: gradient through padding tokens
        dh = _empty(x, M, C)
        ws, rdy = weight_blobs.get(qkv_w, M, C, 3 * C, True)
        call("nmae_linear_bwd_input", dqkv, qkv_w, M, 3 * C, C, rdy, None, dh, ws, device=dev)
        dlw = dlb = None
        if ln_w is not None:
            dx = torch.empty_like(x)
            dlw, dlb = torch.empty_like(ln_w), torch.empty_like(ln_w)
            call("nmae_layernorm_bwd", dh, x, ln_w, mean, rstd, M, C, dout if residual else None, dx, dlw, dlb, device=dev)
        else:
            dx = dh.view_as(x)
            if residual:
                dx = dx + dout
        return dx, dlw, dlb, None, dqw, dqb, dpw, dpb, dtable, None, None, None, None


class MLPFn(Function):
    """[LayerNorm ->] Linear(C,4C) -> GELU(erf) -> Linear(4C,C) [-> x + row_scale * .]; swin_mae3d.py:352-358,367."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, w1, b1, w2, b2, residual, row_scale):
        x = _f32c(x)
        C = x.shape[-1]
        M = x.numel() // C
        T = M // x.shape[0]
        Hd = w1.shape[0]
        dev = x.device
        w1, w2 = _f32c(w1), _f32c(w2)
        if ln_w is not None:
            ln_w, ln_b = _f32c(ln_w), _f32c(ln_b)
            h = torch.empty_like(x)
            mean, rstd = _empty(x, M), _empty(x, M)
            call("nmae_layernorm_fwd", x, ln_w, ln_b, M, C, float(eps), h, mean, rstd, device=dev)
        else:
            h, mean, rstd = x, None, None
        pre, act = _empty(x, M, Hd), _empty(x, M, Hd)
        ws, rdy = weight_blobs.get(w1, M, Hd, C, False)
        call("nmae_linear_fwd", h, w1, b1, M, Hd, C, 1 | rdy, pre, None, None, 1, act, ws, device=dev)
        out = torch.empty_like(x)
        if row_scale is not None:
            row_scale = _f32c(row_scale)
        ws, rdy = weight_blobs.get(w2, M, C, Hd, False)
        call("nmae_linear_fwd", act, w2, b2, M, C, Hd, (2 if residual else 0) | rdy, None, x if residual else None,
             row_scale if residual else None, T, out, ws, device=dev)
        ctx.save_for_backward(x, ln_w, mean, rstd, h if ln_w is not None else None, w1, w2, pre, act, row_scale)
        ctx.cfg = (residual, b1 is not None, b2 is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, ln_w, mean, rstd, h, w1, w2, pre, act, row_scale = ctx.saved_tensors
        residual, has_b1, has_b2 = ctx.cfg
        dout = _f32c(dout)
        C = x.shape[-1]
        M = x.numel() // C
        T = M // x.shape[0]
        Hd = w1.shape[0]
        dev = x.device
        if h is None:
            h = x
        d2 = dout
        if residual and row_scale is not None:
            d2 = torch.empty_like(dout)
            call("nmae_scale_rows", d2, dout, row_scale, T, M, C, device=dev)
        dpre = _empty(x, M, Hd)
        ws, rdy = weight_blobs.get(w2, M, Hd, C, True)
        call("nmae_linear_bwd_input", d2, w2, M, C, Hd, 1 | rdy, pre, dpre, ws, device=dev)      # fused GELU'
        dw2 = torch.empty_like(w2)
        db2 = _empty(x, C) if has_b2 else None
        call("nmae_linear_bwd_weight", d2, act, M, C, Hd, dw2, db2, device=dev)
        dw1 = torch.empty_like(w1)
        db1 = _empty(x, Hd) if has_b1 else None
        call("nmae_linear_bwd_weight", dpre, h, M, Hd, C, dw1, db1, device=dev)
        dh = _empty(x, M, C)
        ws, rdy = weight_blobs.get(w1, M, C, Hd, True)
        call("nmae_linear_bwd_input", dpre, w1, M, Hd, C, rdy, None, dh, ws, device=dev)
        dlw = dlb = None
        if ln_w is not None:
            dx = torch.empty_like(x)
            dlw, dlb = torch.empty_like(ln_w), torch.empty_like(ln_w)
            call("nmae_layernorm_bwd", dh, x, ln_w, mean, rstd, M, C, dout if residual else None, dx, dlw, dlb, device=dev)
        else:
            dx = dh.view_as(x)
            if residual:
                dx = dx + dout
        return dx, dlw, dlb, None, dw1, db1, dw2, db2, None, None


# --------------------------------------------------------------------------------------------- patch merging
class PatchMergeFn(Function):
    """2x2x2 gather + LayerNorm(8C) + Linear(8C->2C, no bias); swin_mae3d.py:372-414."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, red_w):
        x, ln_w, ln_b, red_w = _f32c(x), _f32c(ln_w), _f32c(ln_b), _f32c(red_w)
        lead = x.shape[:-4]
        H, W, D, C = x.shape[-4:]
        B = x.numel() // (H * W * D * C)
        H2, W2, D2 = (H + 1) // 2, (W + 1) // 2, (D + 1) // 2
        rows = B * H2 * W2 * D2
        N = red_w.shape[0]
        if N != 2 * C:
            raise ValueError("PatchMerging: only expand_dim=True (8C -> 2C) is implemented")
        normed = _empty(x, rows, 8 * C)
        mean, rstd = _empty(x, rows), _empty(x, rows)
        out = _empty(x, *lead, H2, W2, D2, N)
        call("nmae_patch_merge_fwd", x, ln_w, ln_b, red_w, B, H, W, D, C, float(eps), normed, mean, rstd, out,
             _empty(x, red_w.numel()), device=x.device)
        ctx.save_for_backward(x, ln_w, red_w, normed, mean, rstd)
        ctx.dims = (B, H, W, D, C)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, ln_w, red_w, normed, mean, rstd = ctx.saved_tensors
        B, H, W, D, C = ctx.dims
        dout = _f32c(dout)
        ws = torch.empty_like(normed)
        dx = torch.empty_like(x)
        dlw, dlb = torch.empty_like(ln_w), torch.empty_like(ln_w)
        drw = torch.empty_like(red_w)
        call("nmae_patch_merge_bwd", dout, x, ln_w, red_w, normed, mean, rstd, B, H, W, D, C, ws, dx, dlw, dlb, drw,
             _empty(x, red_w.numel()), device=x.device)
        return dx, dlw, dlb, None, drw


# --------------------------------------------------------------------------------------------- decoder
class ConvTransposeCatFn(Function):
    """ConvTranspose3d(kernel == stride) written straight into the channel-concat buffer with the skip
    (unetr_block.py:151-158,193-198).  x, skip, result are channels-last (B,X,Y,Z,C)."""

    @staticmethod
    def forward(ctx, x, w, b, skip, k):
        x, w = _f32c(x), _f32c(w)
        B, X, Y, Z, Cin = x.shape
        Cout = w.shape[1]
        Cs = 0 if skip is None else skip.shape[-1]
        ld = Cout + Cs
        out = _empty(x, B, X * k, Y * k, Z * k, ld)
        call("nmae_convT_k_eq_s_fwd", x, w, None if b is None else _f32c(b), B, X, Y, Z, Cin, Cout, k, out, ld,
             _empty(x, w.numel()), device=x.device)
        if skip is not None:
            skip = _f32c(skip)
            if tuple(skip.shape[:4]) != (B, X * k, Y * k, Z * k):
                raise ValueError(f"skip {tuple(skip.shape)} does not match upsampled {(B, X * k, Y * k, Z * k)}")
            rows = B * X * Y * Z * k ** 3
            call("nmae_copy_cols", out.view(-1)[Cout:], ld, skip, Cs, rows, Cs, device=x.device)
        ctx.save_for_backward(x, w)
        ctx.cfg = (k, Cout, Cs, b is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        k, Cout, Cs, has_b = ctx.cfg
        dout = _f32c(dout)
        B, X, Y, Z, Cin = x.shape
        ld = Cout + Cs
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        db = _empty(x, Cout) if has_b else None
        call("nmae_convT_k_eq_s_bwd", dout, ld, x, w, B, X, Y, Z, Cin, Cout, k, dx, dw, db, _empty(x, w.numel()), device=x.device)
        dskip = None
        if Cs:
            dskip = _empty(x, B, X * k, Y * k, Z * k, Cs)
            rows = B * X * Y * Z * k ** 3
            call("nmae_copy_cols", dskip, Cs, dout.view(-1)[Cout:], ld, rows, Cs, device=x.device)
        return dx, dw, db, dskip, None


def conv3_image(x_cl: torch.Tensor, type_dy: bool = False):
    """bf16 hi/lo tensor-core operand image of a channels-last volume (include/nmae.h: nmae_conv3_image_build), or None when
    the channel count does not qualify for the tcgen05 path (multiple of 48)."""
    B, X, Y, Z, C = x_cl.shape
    nbytes = conv3_image_bytes(B, X, Y, Z, C)
    if nbytes == 0:
        return None
    img = torch.empty(nbytes, dtype=torch.uint8, device=x_cl.device)
    call("nmae_conv3_image_build", x_cl, C, 0, B, X, Y, Z, C, 1 if type_dy else 0, img, device=x_cl.device)
    return img


def conv3h_image(x_cl: torch.Tensor, stats: Optional[torch.Tensor] = None, eps: float = 1e-5, slope: float = 0.01,
                 scale: Optional[torch.Tensor] = None):
    """fp16 single-pass operand image of a channels-last volume (include/nmae.h: nmae_conv3h_image_build), optionally of
    LeakyReLU(InstanceNorm(x)) when `stats` is given; None when the channel count is a multiple of neither 48 nor 64."""
    B, X, Y, Z, C = x_cl.shape
    nbytes = conv3h_image_bytes(B, X, Y, Z, C)
    if nbytes == 0:
        return None
    img = torch.empty(nbytes, dtype=torch.uint8, device=x_cl.device)
    call("nmae_conv3h_image_build", x_cl, C, 0, B, X, Y, Z, C, stats, float(eps), float(slope), scale, img, device=x_cl.device)
    return img


class Conv3x3x3Fn(Function):
    """nn.Conv3d(kernel 3, padding 1, stride 1) on a channels-last volume (unetr_block.py:40-56)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _f32c(x), _f32c(w)
        B, X, Y, Z, Cin = x.shape
        Co = w.shape[0]
        wws = _empty(x, 27 * Cin * Co)
        y = _empty(x, B, X, Y, Z, Co)
        ctx.hmode = (_conv_precision == "fp16" and conv3h_image_bytes(B, X, Y, Z, Cin) != 0 and conv3h_image_bytes(B, X, Y, Z, Co) != 0
                     and (Cin % 48 == 0) == (Co % 48 == 0))
        ctx.has_bias = b is not None
        if ctx.hmode:
            ximg = conv3h_image(x)
            call("nmae_conv3h_fwd", ximg, w, None if b is None else _f32c(b), B, X, Y, Z, Cin, Co, wws, y, device=x.device)
            ctx.save_for_backward(x, w, ximg)
            return y
        ximg = conv3_image(x)
        call("nmae_conv3x3x3_fwd", x, ximg, w, None if b is None else _f32c(b), B, X, Y, Z, Cin, Co, wws, y, device=x.device)
        # the operand image is kept for the weight gradient (it replaces x there)
        ctx.save_for_backward(x, w, ximg)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, ximg = ctx.saved_tensors
        dy = _f32c(dy)
        B, X, Y, Z, Cin = x.shape
        Co = w.shape[0]
        wws = _empty(x, 27 * Cin * Co)
        dx = torch.empty_like(x)
        if ctx.hmode:
            # power-of-two scale from the largest |dy| (bookkeeping on 2 floats; the image build applies it)
            amax = dy.abs().amax().clamp_min(1e-30)
            scale = torch.exp2(torch.floor(torch.log2(1024.0 / amax))).reshape(1)
            inv = (1.0 / scale)
            dyimg = conv3h_image(dy, scale=scale)
            call("nmae_conv3h_dgrad", dyimg, inv, w, B, X, Y, Z, Cin, Co, wws, dx, 0, device=x.device)
            dw = torch.empty_like(w)
            call("nmae_conv3h_wgrad", dyimg, inv, ximg, B, X, Y, Z, Cin, Co, dw, device=x.device)
            db = None
            if ctx.has_bias:
                db = _empty(x, Co)
                call("nmae_colsum", dy, B * X * Y * Z, Co, Co, db, device=x.device)
            return dx, dw, db
        dyimg = conv3_image(dy)
        call("nmae_conv3x3x3_dgrad", dy, dyimg, w, B, X, Y, Z, Cin, Co, wws, dx, 0, device=x.device)
        dw = torch.empty_like(w)
        db = _empty(x, Co) if ctx.has_bias else None
        call("nmae_conv3x3x3_wgrad", dy, dyimg, x, ximg, B, X, Y, Z, Cin, Co, wws, dw, db, device=x.device)
        return dx, dw, db


class ResBlockFn(Function):
    """UnetResBlock (unetr_block.py:57-71): conv3^3 -> IN -> LReLU -> conv3^3 -> IN -> (+ IN(conv1^3(x)) | + x) -> LReLU,
    on channels-last volumes.  One autograd node so that the five full-resolution intermediates are
    allocated exactly once."""

    EPS = 1e-5

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, slope, w_out=None, b_out=None):
        """w_out (4, Co) / b_out (4): fuse the 1x1x1 output convolution (UnetOutBlock, unetr_block.py:96-116) - the function then
        returns its (B,X,Y,Z,4) result and the backward never materialises the Co-channel gradient of the block's output."""
        x, w1, b1, w2, b2 = _f32c(x), _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2)
        ctx.fused_out = w_out is not None
        ctx.w_out = _f32c(w_out) if w_out is not None else None
        ctx.b_out = _f32c(b_out) if b_out is not None else None
        B, X, Y, Z, Cin = x.shape
        Co = w1.shape[0]
        V = X * Y * Z
        dev = x.device
        wws = _empty(x, 27 * max(Cin, Co) * Co)
        y1 = _empty(x, B, X, Y, Z, Co)
        st1 = _empty(x, B, Co, 2, dtype=torch.float64)
        # "fp16" precision mode: single-pass fp16 operand images and kernels when both channel counts qualify (48- or 64-groups)
        hmode = (_conv_precision == "fp16" and conv3h_image_bytes(B, X, Y, Z, Cin) != 0 and conv3h_image_bytes(B, X, Y, Z, Co) != 0
                 and (Cin % 48 == 0) == (Co % 48 == 0))
        ctx.hmode = hmode
        if hmode:
            ximg = conv3h_image(x)
            call("nmae_conv3h_fwd", ximg, w1, b1, B, X, Y, Z, Cin, Co, wws, y1, device=dev)
            call("nmae_instnorm_stats", y1, B, V, Co, st1, device=dev)
            a1 = None
            a1img = conv3h_image(y1, st1, ResBlockFn.EPS, slope)
            y2 = torch.empty_like(y1)
            st2 = torch.empty_like(st1)
            call("nmae_conv3h_fwd", a1img, w2, b2, B, X, Y, Z, Co, Co, wws, y2, device=dev)
            call("nmae_instnorm_stats", y2, B, V, Co, st2, device=dev)
            return ResBlockFn._finish_forward(ctx, x, w1, w2, w3, b3, y1, st1, a1, y2, st2, ximg, a1img, slope)
        if ctx.fused_out:
            raise ValueError("ResBlockFn: the fused output convolution needs the fp16 convolution path")
        ximg = conv3_image(x)
        call("nmae_conv3x3x3_fwd", x, ximg, w1, b1, B, X, Y, Z, Cin, Co, wws, y1, device=dev)
        call("nmae_instnorm_stats", y1, B, V, Co, st1, device=dev)
        nbytes = conv3_image_bytes(B, X, Y, Z, Co)
        if nbytes:
            # tensor-core path: LeakyReLU(IN(y1)) is written straight into the operand image of conv2; the fp32 activation
            # is never materialised (the backward recomputes its sign from y1 and the statistics)
            a1 = None
            a1img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            call("nmae_conv3_image_build_in_lrelu", y1, st1, B, X, Y, Z, Co, ResBlockFn.EPS, slope, a1img, device=dev)
        else:
            a1 = torch.empty_like(y1)
            call("nmae_in_lrelu_apply_fwd", y1, st1, None, None, B, V, Co, ResBlockFn.EPS, slope, a1, device=dev)
            a1img = None
        y2 = torch.empty_like(y1)
        st2 = torch.empty_like(st1)
        call("nmae_conv3x3x3_fwd", a1, a1img, w2, b2, B, X, Y, Z, Co, Co, wws, y2, device=dev)
        call("nmae_instnorm_stats", y2, B, V, Co, st2, device=dev)
        return ResBlockFn._finish_forward(ctx, x, w1, w2, w3, b3, y1, st1, a1, y2, st2, ximg, a1img, slope)

    @staticmethod
    def _finish_forward(ctx, x, w1, w2, w3, b3, y1, st1, a1, y2, st2, ximg, a1img, slope):
        B, X, Y, Z, Cin = x.shape
        Co = w1.shape[0]
        V = X * Y * Z
        dev = x.device
        out = torch.empty_like(y1)
        pred = None
        if w3 is not None:
            w3, b3 = _f32c(w3), _f32c(b3)
            y3 = torch.empty_like(y1)
            st3 = torch.empty_like(st1)
            call("nmae_linear_fwd", x, w3, b3, B * V, Co, Cin, 0, None, None, None, 1, y3, _empty(x, w3.numel()), device=dev)
            call("nmae_instnorm_stats", y3, B, V, Co, st3, device=dev)
            call("nmae_in_lrelu_apply_fwd", y2, st2, y3, st3, B, V, Co, ResBlockFn.EPS, slope, out, device=dev)
        else:
            if Cin != Co:
                raise ValueError("identity residual needs in_channels == out_channels")
            y3 = st3 = None
            if ctx.fused_out and ctx.w_out.shape[0] == 4 and Co <= 128:
                # the output convolution is evaluated while the block's result is on chip
                pred = _empty(x, B, X, Y, Z, 4)
                call("nmae_in_lrelu_apply_out_fwd", y2, st2, x, None, B, V, Co, ResBlockFn.EPS, slope, out, ctx.w_out, ctx.b_out, pred,
                     device=dev)
            else:
                call("nmae_in_lrelu_apply_fwd", y2, st2, x, None, B, V, Co, ResBlockFn.EPS, slope, out, device=dev)
        # the operand images of x and a1 are kept: the weight gradients read them instead of the fp32 volumes
        ctx.save_for_backward(x, w1, w2, w3, y1, st1, a1, y2, st2, y3, st3, out, ximg, a1img)
        ctx.slope = slope
        if ctx.fused_out:
            if pred is None:
                pred = _empty(x, B, X, Y, Z, ctx.w_out.shape[0])
                call("nmae_linear_fwd", out, ctx.w_out, ctx.b_out, B * V, ctx.w_out.shape[0], Co, 0, None, None, None, 1, pred, None,
                     device=dev)
            return pred
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, w1, w2, w3, y1, st1, a1, y2, st2, y3, st3, out, ximg, a1img = ctx.saved_tensors
        slope = ctx.slope
        dout = _f32c(dout)
        B, X, Y, Z, Cin = x.shape
        Co = w1.shape[0]
        V = X * Y * Z
        dev = x.device
        eps = ResBlockFn.EPS
        wws = _empty(x, 27 * max(Cin, Co) * Co)
        sums = _empty(x, B, Co, 3, dtype=torch.float64)
        dx = torch.empty_like(x)
        # tensor-core path: the gradients wrt the convolution outputs (dy2, dy1) are only ever read by the dgrad / weight-gradient
        # kernels, so the InstanceNorm backward writes them straight into operand images and no fp32 copy exists
        nbytes = conv3_image_bytes(B, X, Y, Z, Co)
        if ctx.hmode:
            return ResBlockFn._backward_h(ctx, dout, x, w1, w2, w3, y1, st1, y2, st2, y3, st3, out, ximg, a1img, slope)
        # (the fused output convolution exists on the fp16 path only)
        tc2 = a1img is not None and nbytes != 0
        tc1 = tc2 and ximg is not None
        # the bias gradients of conv1/conv2 are the column sums of dy1/dy2: accumulated by the kernel that writes them
        dw2, db2 = torch.empty_like(w2), _empty(x, Co)
        dy3 = torch.empty_like(y2) if w3 is not None else None
        dres = dx if w3 is None else None
        if tc2:
            dy2 = None
            dy2img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            call("nmae_in_lrelu_apply_bwd_image", dout, out, y2, st2, y3, st3, B, X, Y, Z, Co, eps, slope, sums, dy2img, dy3, dres,
                 db2, None, device=dev)
        else:
            dy2 = torch.empty_like(y2)
            call("nmae_in_lrelu_apply_bwd", dout, out, y2, st2, y3, st3, B, V, Co, eps, slope, sums, dy2, dy3, dres, db2, None,
                 device=dev)
            dy2img = conv3_image(dy2)
        call("nmae_conv3x3x3_wgrad", dy2, dy2img, a1, a1img, B, X, Y, Z, Co, Co, wws, dw2, None, device=dev)
        da1 = torch.empty_like(y1)
        call("nmae_conv3x3x3_dgrad", dy2, dy2img, w2, B, X, Y, Z, Co, Co, wws, da1, 0, device=dev)
        del dy2, dy2img
        dw1, db1 = torch.empty_like(w1), _empty(x, Co)
        # a1 is None on the tensor-core path: LeakyReLU(IN(y1)) has the sign of IN(y1), which the kernel recomputes
        if tc1:
            dy1 = None
            dy1img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            call("nmae_in_lrelu_apply_bwd_image", da1, a1, y1, st1, None, None, B, X, Y, Z, Co, eps, slope, sums, dy1img, None, None,
                 db1, None, device=dev)
        else:
            dy1 = torch.empty_like(y1)
            call("nmae_in_lrelu_apply_bwd", da1, a1, y1, st1, None, None, B, V, Co, eps, slope, sums, dy1, None, None, db1, None,
                 device=dev)
            dy1img = conv3_image(dy1)
        del da1
        call("nmae_conv3x3x3_wgrad", dy1, dy1img, x, ximg, B, X, Y, Z, Cin, Co, wws, dw1, None, device=dev)
        # identity residual: dx already holds its gradient -> accumulate the conv1 dgrad on top
        call("nmae_conv3x3x3_dgrad", dy1, dy1img, w1, B, X, Y, Z, Cin, Co, wws, dx, 0 if w3 is not None else 1, device=dev)
        del dy1img
        dw3 = db3 = None
        if w3 is not None:
            dw3, db3 = torch.empty_like(w3), _empty(x, Co)
            call("nmae_linear_bwd_weight", dy3, x, B * V, Co, Cin, dw3, db3, device=dev)
            call("nmae_linear_bwd_input", dy3, w3, B * V, Co, Cin, 4, None, dx, _empty(dy3, w3.numel()), device=dev)
        return dx, dw1, db1, dw2, db2, dw3, db3, None


def _resblock_backward_h(ctx, dout, x, w1, w2, w3, y1, st1, y2, st2, y3, st3, out, ximg, a1img, slope):
    """Backward of ResBlockFn in the "fp16" precision mode: gradients wrt the convolution outputs are written straight into scaled
    fp16 operand images (one device float per image holds the reciprocal scale) that the dgrad / weight-gradient kernels read."""
    B, X, Y, Z, Cin = x.shape
    Co = w1.shape[0]
    V = X * Y * Z
    dev = x.device
    eps = ResBlockFn.EPS
    wws = _empty(x, 27 * max(Cin, Co) * Co)
    sums = _empty(x, workspace_bytes("nmae_in_lrelu_bwd_sums_ws_bytes", B, Co) // 8, dtype=torch.float64)
    scal = _empty(x, 4)                      # [amax scratch, inv_scale(dy2), amax scratch, inv_scale(dy1)]
    nbytes = conv3h_image_bytes(B, X, Y, Z, Co)
    dx = torch.empty_like(x)
    dw2, db2 = torch.empty_like(w2), _empty(x, Co)
    dy3 = torch.empty_like(y2) if w3 is not None else None
    dres = dx if w3 is None else None
    dy2img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dw_out = db_out = dp4 = None
    if ctx.fused_out:
        # dout is the gradient wrt the (B,X,Y,Z,4) result of the fused output convolution: its weight gradient reads `out`; the
        # Co-channel gradient dout @ w_out is evaluated inside the InstanceNorm backward kernels instead of being written
        dp4, n_out = dout, ctx.w_out.shape[0]
        dw_out, db_out = torch.empty_like(ctx.w_out), _empty(x, n_out)
        # the output convolution's weight gradient comes out of the same reduction pass when the block has no shortcut convolution
        fuse_wg = w3 is None and n_out == 4
        if not fuse_wg:
            call("nmae_linear_bwd_weight", dp4, out, B * V, n_out, Co, dw_out, db_out, device=dev)
        dout = None
    else:
        fuse_wg = False
    call("nmae_in_lrelu_apply_bwd_image_h", dout, out, y2, st2, y3, st3, B, X, Y, Z, Co, eps, slope, sums, scal[0:1], dy2img, scal[1:2],
         dy3, dres, db2, None, dp4, ctx.w_out if ctx.fused_out else None, dw_out if fuse_wg else None, db_out if fuse_wg else None,
         device=dev)
    call("nmae_conv3h_wgrad", dy2img, scal[1:2], a1img, B, X, Y, Z, Co, Co, dw2, device=dev)
    da1 = torch.empty_like(y1)
    call("nmae_conv3h_dgrad", dy2img, scal[1:2], w2, B, X, Y, Z, Co, Co, wws, da1, 0, device=dev)
    del dy2img
    dw1, db1 = torch.empty_like(w1), _empty(x, Co)
    dy1img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    call("nmae_in_lrelu_apply_bwd_image_h", da1, None, y1, st1, None, None, B, X, Y, Z, Co, eps, slope, sums, scal[2:3], dy1img, scal[3:4],
         None, None, db1, None, None, None, None, None, device=dev)
    del da1
    call("nmae_conv3h_wgrad", dy1img, scal[3:4], ximg, B, X, Y, Z, Cin, Co, dw1, device=dev)
    call("nmae_conv3h_dgrad", dy1img, scal[3:4], w1, B, X, Y, Z, Cin, Co, wws, dx, 0 if w3 is not None else 1, device=dev)
    del dy1img
    dw3 = db3 = None
    if w3 is not None:
        dw3, db3 = torch.empty_like(w3), _empty(x, Co)
        call("nmae_linear_bwd_weight", dy3, x, B * V, Co, Cin, dw3, db3, device=dev)
        call("nmae_linear_bwd_input", dy3, w3, B * V, Co, Cin, 4, None, dx, _empty(dy3, w3.numel()), device=dev)
    return dx, dw1, db1, dw2, db2, dw3, db3, None, dw_out, db_out


ResBlockFn._backward_h = staticmethod(_resblock_backward_h)


class InstNormLReLUFn(Function):
    """nn.InstanceNorm3d (no affine, eps 1e-5) + nn.LeakyReLU(slope) on a channels-last volume: the conv -> IN -> LeakyReLU(0.2)
    stages of the legacy decoder (swin_mae3d.py:593-610)."""

    @staticmethod
    def forward(ctx, x, slope, eps):
        x = _f32c(x)
        B, X, Y, Z, C = x.shape
        V = X * Y * Z
        st = _empty(x, B, C, 2, dtype=torch.float64)
        call("nmae_instnorm_stats", x, B, V, C, st, device=x.device)
        out = torch.empty_like(x)
        call("nmae_in_lrelu_apply_fwd", x, st, None, None, B, V, C, float(eps), float(slope), out, device=x.device)
        ctx.save_for_backward(x, st)
        ctx.cfg = (float(slope), float(eps))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, st = ctx.saved_tensors
        slope, eps = ctx.cfg
        dout = _f32c(dout)
        B, X, Y, Z, C = x.shape
        sums = _empty(x, B, C, 3, dtype=torch.float64)
        dx = torch.empty_like(x)
        call("nmae_in_lrelu_apply_bwd", dout, None, x, st, None, None, B, X * Y * Z, C, eps, slope, sums, dx, None, None, None, None,
             device=x.device)
        return dx, None, None


class UpsampleTrilinearFn(Function):
    """nn.Upsample(size=size, mode="trilinear", align_corners=False) on a channels-last volume (swin_mae3d.py:597,602,607)."""

    @staticmethod
    def forward(ctx, x, size):
        x = _f32c(x)
        B, X, Y, Z, C = x.shape
        Xf, Yf, Zf = (int(v) for v in size)
        out = _empty(x, B, Xf, Yf, Zf, C)
        call("nmae_upsample_trilinear_fwd", x, out, B, X, Y, Z, Xf, Yf, Zf, C, device=x.device)
        ctx.dims = (B, X, Y, Z, Xf, Yf, Zf, C)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, X, Y, Z, Xf, Yf, Zf, C = ctx.dims
        dout = _f32c(dout)
        dx = _empty(dout, B, X, Y, Z, C)
        call("nmae_upsample_trilinear_bwd", dout, dx, B, X, Y, Z, Xf, Yf, Zf, C, device=dout.device)
        return dx, None


class MAELossFn(Function):
    """forward_loss (swin_mae3d.py:1513-1563) -> tensor [loss, loss_rgb, loss_alpha]."""

    @staticmethod
    def forward(ctx, pred, x, ext, tok_mask, p):
        pred, x = _f32c(pred), _f32c(x)
        B, R = x.shape[0], x.shape[2]
        if tuple(pred.shape) != (B, R, R, R, 4):
            raise ValueError(f"pred must be channels-last (B,R,R,R,4), got {tuple(pred.shape)}")
        sums = _empty(x, 4, dtype=torch.float64)
        out3 = _empty(x, 3)
        call("nmae_mae_loss_fwd", x, pred, ext, tok_mask, B, R, p, sums, out3, device=x.device)
        ctx.save_for_backward(pred, x, ext, tok_mask, sums)
        ctx.p = p
        return out3

    @staticmethod
    @once_differentiable
    def backward(ctx, g3):
        pred, x, ext, tok_mask, sums = ctx.saved_tensors
        B, R = x.shape[0], x.shape[2]
        dpred = torch.empty_like(pred)
        call("nmae_mae_loss_bwd", x, pred, ext, tok_mask, B, R, ctx.p, sums, _f32c(g3), dpred, device=x.device)
        return dpred, None, None, None, None


# --------------------------------------------------------------------------------------------- thin wrappers
def layer_norm(x, w, b, eps=1e-5):
    return LayerNormFn.apply(x, w, b, eps)


def linear(x, w, b=None):
    return LinearFn.apply(x, w, b)


def window_attention(x, qkv_w, qkv_b, proj_w, proj_b, table, num_heads, shift, ln_w=None, ln_b=None, eps=1e-5,
                     residual=False, row_scale: Optional[torch.Tensor] = None):
    return WindowAttentionFn.apply(x, ln_w, ln_b, eps, qkv_w, qkv_b, proj_w, proj_b, table, num_heads, shift, residual, row_scale)


def mlp(x, w1, b1, w2, b2, ln_w=None, ln_b=None, eps=1e-5, residual=False, row_scale: Optional[torch.Tensor] = None):
    return MLPFn.apply(x, ln_w, ln_b, eps, w1, b1, w2, b2, residual, row_scale)


def pad_grids(grids, R: int):
    """torch_utils.py:56-90 + swin_mae3d.py:1432-1448: list of (4,X,Y,Z) -> (B,4,R,R,R) and the (B,3) int32
    extents that stand in for the reference's dense pad mask."""
    dev = grids[0].device
    B = len(grids)
    batch = torch.empty(B, 4, R, R, R, dtype=torch.float32, device=dev)
    ext = []
    for b, g in enumerate(grids):
        g = _f32c(g)
        if g.dim() != 4 or g.shape[0] != 4:
            raise ValueError(f"expected (4,X,Y,Z) grids, got {tuple(g.shape)}")
        _, X, Y, Z = g.shape
        call("nmae_pad_grid", g, X, Y, Z, batch, b, R, device=dev)
        ext.append([min(X, R), min(Y, R), min(Z, R)])     # oversize scenes are cropped, as F.pad's negative pads do
    return batch, torch.tensor(ext, dtype=torch.int32).to(dev, non_blocking=True)


def ingest_scenes(raw, R: int, normalize_density: bool = True, aug=None):
    """Raw scene arrays straight from the .npz files -> padded batch, on the GPU (include/nmae.h: nmae_ingest_scene).

    raw: list of CUDA tensors (W,L,H,4), float32 or uint8 (the `rgbsigma` arrays as stored, copied to the device as they are);
    aug: optional list of (rotate, flip_axis1, flip_axis2) booleans per scene, drawn on the host exactly like
    nerf_rpn/datasets.py:172-234 does (see run_swin_mae3d.draw_augmentation).  Returns (batch (B,4,R,R,R), extents (B,3) int32):
    the pair SwinTransformer_MAE3D_New.transform() yields, so forward_padded() consumes it directly.  Replaces the reference's
    CPU work per scene (np.exp over the whole density channel, transpose copy, flips) and pad_tensor."""
    dev = raw[0].device
    B = len(raw)
    batch = torch.empty(B, 4, R, R, R, dtype=torch.float32, device=dev)
    ext = []
    for b, g in enumerate(raw):
        if g.dim() != 4 or g.shape[3] != 4 or g.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f"expected raw (W,L,H,4) float32 or uint8 scenes, got {tuple(g.shape)} {g.dtype}")
        g = g.contiguous()
        W, L, H, _ = g.shape
        rot, f1, f2 = (bool(v) for v in aug[b]) if aug is not None else (False, False, False)
        call("nmae_ingest_scene", g, int(g.dtype == torch.uint8), int(bool(normalize_density)), W, L, H, int(rot), int(f1), int(f2),
             batch, b, R, device=dev)
        ext.append([min(L if rot else W, R), min(W if rot else L, R), min(H, R)])
    return batch, torch.tensor(ext, dtype=torch.int32).to(dev, non_blocking=True)
