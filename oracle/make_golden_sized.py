"""Known-answer tests at the sizes BASELINE.json names, recorded from the LIVE reference (/root/reference).
Test infrastructure; run in the build container only:

    python oracle/make_golden_sized.py [swin_s160] [swin_t160] [swin_b256] [bench]

Writes tests/golden/kat_sized.json (scalars, per-tensor gradient fingerprints) and tests/golden/kat_sized.npz (sampled
predictions, packed mask bits).  CPU fp32, torch 2.11, reference imported with the numpy.float shim.

Cases (reference: nerf_mae/model/mae/swin_mae3d.py:1571-1599 forward, run_swin_mae3d.py:650-669 train step):
  swin_s160 / swin_t160 : eval forward on one cubic 160^3 grid (A) and on two ragged grids (B); train-mode forward+backward
                          (stochastic depth 0) on the cubic grid -> loss + sum / sum-of-squares of every parameter gradient.
  swin_b256             : BASELINE config 4.  The reference cannot construct swin_b; the convention of SURVEY 8c is
                          applied to the ORACLE side here: heads [4,8,16,32] and the sincos table zero-padded 126 -> 128 channels
                          (get_3d_sincos_pos_embed monkey-patched).  Forward only (a 256^3 backward does not fit this box).
"""
import json
import os
import random
import sys
import time

import numpy
import numpy as np
import torch

numpy.float = float  # shim, torch_utils.py:42
sys.path.insert(0, "/root/reference")
from nerf_mae.model.mae import swin_mae3d as R  # noqa: E402
from nerf_mae.model.mae import torch_utils as RT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
N_SAMPLE = 4096

CFG = {
    "swin_s160": dict(embed_dim=96, depths=[2, 2, 18, 2], heads=[3, 6, 12, 24], res=160),
    "swin_t160": dict(embed_dim=96, depths=[2, 2, 6, 2], heads=[3, 6, 12, 24], res=160),
    "swin_b256": dict(embed_dim=128, depths=[2, 2, 18, 2], heads=[4, 8, 16, 32], res=256),
}


def grids_for(res):
    """The seeded inputs of every sized KAT (the GPU tests regenerate them with the same CPU generator)."""
    g = torch.Generator().manual_seed(1234 + res)
    cubic = torch.rand(4, res, res, res, generator=g)
    ra = torch.rand(4, res - 23, res, res - 60, generator=g)
    rb = torch.rand(4, res, res // 2 + 3, res - 1, generator=g)
    return cubic, ra, rb


def build(name, **kw):
    c = CFG[name]
    if c["embed_dim"] % 6 != 0:
        # SURVEY 8c convention for swin_b: per-axis sincos width 2*floor(C/6)*... = 126 channels, zero tail up to C
        orig = RT.get_3d_sincos_pos_embed

        def padded(embed_dim, grid_size, cls_token=False):
            d = 3 * (embed_dim // 3)
            d -= d % 6
            e = orig(d, grid_size, cls_token)
            return np.concatenate([e, np.zeros(e.shape[:-1] + (embed_dim - e.shape[-1],), e.dtype)], axis=-1)
        R.get_3d_sincos_pos_embed = padded
    torch.manual_seed(0)
    random.seed(0)
    return R.SwinTransformer_MAE3D_New([4, 4, 4], c["embed_dim"], c["depths"], c["heads"], [4, 4, 4], resolution=c["res"],
                                       masking_prob=0.75, **kw)


def run_case(name, kat, arrs, with_grad=True, with_ragged=True):
    c = CFG[name]
    res = c["res"]
    t0 = time.time()
    m = build(name).eval()
    cubic, ra, rb = grids_for(res)
    n_tok = res // 4
    idx = torch.randint(0, n_tok ** 3 * 64 * 4, (N_SAMPLE,), generator=torch.Generator().manual_seed(5))
    arrs[f"{name}.sample_idx"] = idx.numpy()
    out = {}
    cases = [("A", [cubic])] + ([("B", [ra, rb])] if with_ragged else [])
    for tag, grids in cases:
        random.seed(42)
        with torch.no_grad():
            loss, lr, la, pred, valid, target = m(grids, is_eval=True)
        out[tag] = dict(loss=float(loss), loss_rgb=float(lr), loss_alpha=float(la),
                        pred_sq_sum=float((pred.double() ** 2).sum()), pred_sum=float(pred.double().sum()),
                        valid_sum=int(valid.sum()), target_sum=float(target.double().sum()),
                        shapes=[list(pred.shape), list(valid.shape), list(target.shape)])
        for b in range(len(grids)):
            arrs[f"{name}.{tag}.pred_sample{b}"] = pred[b].flatten()[idx].numpy()
        del pred, valid, target
        print(name, tag, out[tag]["loss"], f"{time.time() - t0:.0f}s", flush=True)
    # the mask the forward drew under random.seed(42) (bit-exact requirement)
    random.seed(42)
    _, mk = m.window_masking_3d(torch.zeros(1, n_tok, n_tok, n_tok, 1), p_remove=0.75, mask_token=None)
    arrs[f"{name}.mask42"] = np.packbits(mk[0, ..., 0].numpy().astype(np.uint8))
    if with_grad:
        m2 = build(name, stochastic_depth_prob=0.0)
        m2.load_state_dict(m.state_dict())
        m2.train()
        random.seed(42)
        loss, _, _ = m2([cubic])
        loss.backward()
        out["grad_A_loss"] = float(loss)
        out["grad_A"] = {k: [float(p.grad.double().sum()), float((p.grad.double() ** 2).sum())]
                         for k, p in m2.named_parameters() if p.grad is not None}
        out["grad_A_total_norm"] = float(sum(v[1] for v in out["grad_A"].values()) ** 0.5)
        print(name, "grad", out["grad_A_loss"], out["grad_A_total_norm"], f"{time.time() - t0:.0f}s", flush=True)
    kat[name] = out


def bench_case(kat):
    """The batch bench.py times on rank 0 (swin_s, four 160^3 grids from Generator(0)): eval forward under random.seed(42).
    bench.py repeats it on the GPU before timing and asserts the loss triple (reference swin_mae3d.py:1571-1599)."""
    t0 = time.time()
    m = build("swin_s160").eval()
    gen = torch.Generator().manual_seed(0)
    grids = [torch.rand(4, 160, 160, 160, generator=gen) for _ in range(4)]
    random.seed(42)
    with torch.no_grad():
        loss, lr, la, pred, valid, target = m(grids, is_eval=True)
    kat["bench_swin_s160_b4"] = dict(loss=float(loss), loss_rgb=float(lr), loss_alpha=float(la), valid_sum=int(valid.sum()),
                                     pred_sq_sum=float((pred.double() ** 2).sum()))
    print("bench batch", kat["bench_swin_s160_b4"], f"{time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["swin_s160", "swin_t160"]
    jpath, npath = os.path.join(OUT, "kat_sized.json"), os.path.join(OUT, "kat_sized.npz")
    kat = json.load(open(jpath)) if os.path.exists(jpath) else {}
    arrs = dict(np.load(npath)) if os.path.exists(npath) else {}
    for name in which:
        if name == "bench":
            bench_case(kat)
        else:
            big = name == "swin_b256"
            run_case(name, kat, arrs, with_grad=not big, with_ragged=not big)
        with open(jpath, "w") as f:
            json.dump(kat, f, indent=1)
        np.savez_compressed(npath, **arrs)
    print("wrote", jpath, npath)
