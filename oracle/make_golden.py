"""Generate tests/golden/* from the LIVE reference (/root/reference).  Test infrastructure.

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/make_golden.py
The reference imports with a one-line shim (numpy.float was removed in numpy 2,
SURVEY 0.3-4).  Everything here is CPU fp32, torch 2.11.
"""
import json
import os
import random
import sys

import numpy
import numpy as np
import torch

numpy.float = float  # shim, torch_utils.py:42
sys.path.insert(0, "/root/reference")
from nerf_mae.model.mae import swin_mae3d as R  # noqa: E402
from nerf_mae.model.mae import unetr_block as RU  # noqa: E402
from nerf_mae.model.mae import torch_utils as RT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ops = {}


def put(name, t):
    ops[name] = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def attention_cases():
    g = torch.Generator().manual_seed(7)
    cases = [  # name, (B,H,W,D), C, heads, shift
        ("plain", (2, 8, 8, 8), 64, 2, 0),
        ("shift", (2, 8, 8, 8), 64, 2, 2),
        ("pad_shift", (1, 10, 10, 10), 96, 3, 2),      # stage-3 geometry: 10 -> 12
        ("pad5_shift", (2, 5, 5, 5), 64, 2, 2),         # stage-4 geometry: 5 -> 8
        ("ragged_shift", (1, 6, 10, 5), 32, 1, 2),
        ("tiny_noshift", (1, 2, 2, 2), 64, 2, 2),       # 2 -> 4: shift auto-disabled (S:68-75)
    ]
    for name, (B, H, W, D), C, nh, sh in cases:
        x = torch.randn(B, H, W, D, C, generator=g)
        qw, qb = torch.randn(3 * C, C, generator=g) * 0.08, torch.randn(3 * C, generator=g) * 0.1
        pw, pb = torch.randn(C, C, generator=g) * 0.08, torch.randn(C, generator=g) * 0.1
        table = torch.randn(343, nh, generator=g) * 0.5
        mod = R.ShiftedWindowAttention(C, [4, 4, 4], [sh] * 3, nh)
        with torch.no_grad():
            mod.qkv.weight.copy_(qw); mod.qkv.bias.copy_(qb)
            mod.proj.weight.copy_(pw); mod.proj.bias.copy_(pb)
            mod.relative_position_bias_table.copy_(table)
            y = mod(x)
        for k, v in dict(x=x, qw=qw, qb=qb, pw=pw, pb=pb, table=table, y=y).items():
            put(f"attn.{name}.{k}", v)
        put(f"attn.{name}.meta", np.array([nh, sh]))
    put("attn.rel_index", R.ShiftedWindowAttention(32, [4, 4, 4], [0, 0, 0], 1).relative_position_index)


def merge_cases():
    g = torch.Generator().manual_seed(8)
    for name, (B, H, W, D), C in [("even", (2, 4, 4, 4), 16), ("odd", (1, 5, 5, 5), 16), ("ragged", (1, 3, 4, 5), 8)]:
        mod = R.PatchMerging(C)
        x = torch.randn(B, H, W, D, C, generator=g)
        with torch.no_grad():
            mod.norm.weight.copy_(torch.randn(8 * C, generator=g)); mod.norm.bias.copy_(torch.randn(8 * C, generator=g))
            mod.reduction.weight.copy_(torch.randn(2 * C, 8 * C, generator=g) * 0.1)
            y = mod(x)
        for k, v in dict(x=x, nw=mod.norm.weight, nb=mod.norm.bias, rw=mod.reduction.weight, y=y).items():
            put(f"merge.{name}.{k}", v)


def block_case():
    torch.manual_seed(9)
    blk = R.SwinTransformerBlock(32, 1, [4, 4, 4], [2, 2, 2], stochastic_depth_prob=0.0,
                                 norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-5)).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn_like(p) * 0.2)
        x = torch.randn(1, 6, 6, 6, 32)
        y = blk(x)
    put("block.x", x); put("block.y", y)
    for k, v in blk.state_dict().items():
        put("block.sd." + k, v)


def decoder_cases():
    torch.manual_seed(10)
    up = RU.UnetrUpBlock(16, 8, 3, 2, res_block=True).eval()
    x, skip = torch.randn(2, 16, 3, 3, 3), torch.randn(2, 8, 6, 6, 6)
    with torch.no_grad():
        y = up(x, skip)
    put("up.skip.x", x); put("up.skip.skip", skip); put("up.skip.y", y)
    for k, v in up.state_dict().items():
        put("up.skip.sd." + k, v)
    up = RU.UnetrUpBlock(8, 4, 3, 4, res_block=True, use_skip=False).eval()
    x = torch.randn(1, 8, 2, 2, 2)
    with torch.no_grad():
        y = up(x)
    put("up.noskip.x", x); put("up.noskip.y", y)
    for k, v in up.state_dict().items():
        put("up.noskip.sd." + k, v)
    out = RU.UnetOutBlock(4, 4)
    x = torch.randn(1, 4, 3, 3, 3)
    with torch.no_grad():
        put("outblock.x", x); put("outblock.y", out(x)); put("outblock.w", out.conv.weight); put("outblock.b", out.conv.bias)


def misc_cases():
    put("pos_embed.96.5", torch.from_numpy(RT.get_3d_sincos_pos_embed(96, 5)).float())
    put("pos_embed.192.3", torch.from_numpy(RT.get_3d_sincos_pos_embed(192, 3)).float())
    t = torch.arange(4 * 3 * 5 * 2, dtype=torch.float32).reshape(4, 3, 5, 2)
    p, m = RT.pad_tensor(t, [8, 8, 8], 0)
    put("pad.in", t); put("pad.out", p); put("pad.mask", m)


def model_kats():
    kat = {}
    torch.manual_seed(0); random.seed(0)
    m = R.SwinTransformer_MAE3D_New([4, 4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [4, 4, 4],
                                    resolution=64, masking_prob=0.75).eval()
    # fingerprint of the seed-0 initial weights (the product model must reproduce the init stream)
    kat["init_fingerprint"] = {k: [float(v.double().sum()), float(v.double().abs().sum())]
                               for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    g = torch.Generator().manual_seed(1234)
    x1 = torch.rand(4, 64, 64, 64, generator=g)
    xa = torch.rand(4, 50, 60, 64, generator=g)
    xb = torch.rand(4, 64, 33, 47, generator=g)
    idx = torch.randint(0, 16 ** 3 * 64 * 4, (4096,), generator=torch.Generator().manual_seed(5))
    for name, grids in (("A", [x1]), ("B", [xa, xb])):
        random.seed(42)
        with torch.no_grad():
            loss, lr, la, pred, valid, target = m(grids, is_eval=True)
        kat[name] = dict(loss=float(loss), loss_rgb=float(lr), loss_alpha=float(la),
                         pred_sum=float(pred.double().sum()), pred_abs_sum=float(pred.double().abs().sum()),
                         pred_sq_sum=float((pred.double() ** 2).sum()),
                         valid_sum=int(valid.sum()), target_sum=float(target.double().sum()),
                         shapes=[list(pred.shape), list(valid.shape), list(target.shape)])
        put(f"kat.{name}.pred_sample", pred[0].flatten()[idx])
    put("kat.sample_idx", idx)
    # gradient KAT (train-mode semantics but stochastic depth off): grads of a few tensors
    m2 = R.SwinTransformer_MAE3D_New([4, 4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [4, 4, 4],
                                     resolution=64, masking_prob=0.75, stochastic_depth_prob=0.0)
    m2.load_state_dict(m.state_dict()); m2.train()
    random.seed(42)
    loss, _, _ = m2([x1])
    loss.backward()
    kat["grad_A"] = {k: [float(p.grad.double().sum()), float((p.grad.double() ** 2).sum())]
                     for k, p in m2.named_parameters() if p.grad is not None}
    kat["grad_A_loss"] = float(loss)
    # mask KAT: seed 123, 1000 draws < 0.75 == mask[0,::4,::4,::4,0]
    for n_tok, seed in ((40, 123), (16, 42), (10, 3)):
        random.seed(seed)
        _, mk = m.window_masking_3d(torch.zeros(1, n_tok, n_tok, n_tok, 2), p_remove=0.75, mask_token=None)
        put(f"mask.{n_tok}.{seed}", np.packbits(mk[0, ..., 0].numpy().astype(np.uint8)))
    with open(os.path.join(OUT, "kat_model.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    attention_cases(); merge_cases(); block_case(); decoder_cases(); misc_cases(); model_kats()
    np.savez_compressed(os.path.join(OUT, "golden_ops.npz"), **ops)
    print("wrote", len(ops), "arrays;", os.path.getsize(os.path.join(OUT, "golden_ops.npz")) / 1e6, "MB")
