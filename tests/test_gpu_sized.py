"""GPU parity at the sizes BASELINE.json names, against known answers recorded from the LIVE reference
(oracle/make_golden_sized.py -> tests/golden/kat_sized.{json,npz}; reference swin_mae3d.py:1571-1599, run_swin_mae3d.py:650-669).

swin_s and swin_t at 160^3 (configs 2 and 3): eval forward on one cubic grid and on two ragged grids (loss triple, 4096 sampled
predictions per grid, |pred|^2, valid count, mask bits), and the gradients of a train-mode step (per-tensor sum of squares).
swin_b at 256^3 (config 4) forward only, under the SURVEY 8c convention (the reference cannot construct swin_b).

Tolerance: north_star 1e-3 relative on loss / predictions; mask and valid counts bit-exact.  Every case runs in each convolution
precision mode of the library ("bf16x3": fp32-class three-pass operands; "fp16": single-pass fp16 operands, the TF32-class mode).
"""
import json
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODEL_TOL = 1e-3
GRAD_TOL = 4e-2      # per-tensor |grad|^2 (LeakyReLU kinks: see tests/test_gpu_parity.py rel_trim)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def N():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import nerf_mae_b200
    nerf_mae_b200.lib()
    return nerf_mae_b200


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat_sized.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def arrs():
    return dict(np.load(os.path.join(GOLDEN, "kat_sized.npz")))


def grids_for(res):
    """Must stay identical to oracle/make_golden_sized.py:grids_for."""
    g = torch.Generator().manual_seed(1234 + res)
    cubic = torch.rand(4, res, res, res, generator=g)
    ra = torch.rand(4, res - 23, res, res - 60, generator=g)
    rb = torch.rand(4, res, res // 2 + 3, res - 1, generator=g)
    return cubic, ra, rb


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


NAMES = {"swin_s160": ("swin_s", 160), "swin_t160": ("swin_t", 160), "swin_b256": ("swin_b", 256)}


def _model(N, name, mode, **kw):
    backbone, res = NAMES[name]
    torch.manual_seed(0)
    random.seed(0)
    m = N.build_model(backbone, res, 0.75, **kw)
    N.set_conv_precision(mode)
    return m, res


@pytest.mark.parametrize("mode", ["bf16x3", "fp16"])
@pytest.mark.parametrize("name,case", [("swin_s160", "A"), ("swin_s160", "B"), ("swin_t160", "A"), ("swin_t160", "B"),
                                       ("swin_b256", "A")])
def test_sized_forward_kat(N, kat, arrs, name, case, mode):
    if name not in kat:
        pytest.skip(f"{name} not recorded in kat_sized.json")
    try:
        m, res = _model(N, name, mode)
        m = m.cuda().eval()
        cubic, ra, rb = grids_for(res)
        grids = [cubic.cuda()] if case == "A" else [ra.cuda(), rb.cuda()]
        del cubic, ra, rb
        random.seed(42)
        with torch.no_grad():
            loss, lr, la, pred, valid, target = m(grids, is_eval=True)
        k = kat[name][case]
        assert [list(pred.shape), list(valid.shape), list(target.shape)] == k["shapes"]
        assert int(valid.sum()) == k["valid_sum"]                                            # bit-exact
        mask = m._tok_mask_u8.cpu().numpy().astype(np.uint8)
        assert np.array_equal(np.packbits(mask), arrs[f"{name}.mask42"])                      # bit-exact
        for got, key in ((loss, "loss"), (lr, "loss_rgb"), (la, "loss_alpha")):
            assert abs(float(got) - k[key]) <= MODEL_TOL * abs(k[key]), (key, float(got), k[key])
        idx = torch.from_numpy(arrs[f"{name}.sample_idx"]).cuda()
        for b in range(len(grids)):
            want = torch.from_numpy(arrs[f"{name}.{case}.pred_sample{b}"])
            assert rel(pred[b].flatten()[idx], want) < MODEL_TOL, (b, rel(pred[b].flatten()[idx], want))
        assert abs(float((pred.double() ** 2).sum()) - k["pred_sq_sum"]) <= MODEL_TOL * k["pred_sq_sum"]
    finally:
        N.set_conv_precision(None)
        torch.cuda.empty_cache()


@pytest.mark.parametrize("mode", ["bf16x3", "fp16"])
@pytest.mark.parametrize("name", ["swin_s160", "swin_t160"])
def test_sized_gradient_kat(N, kat, name, mode):
    """Train-mode step (stochastic depth 0) on the cubic grid: loss and every parameter gradient's sum of squares."""
    try:
        m, res = _model(N, name, mode, stochastic_depth_prob=0.0)
        m = m.cuda().train()
        cubic, _, _ = grids_for(res)
        random.seed(42)
        loss, _, _ = m([cubic.cuda()])
        loss.backward()
        k = kat[name]
        assert abs(float(loss) - k["grad_A_loss"]) <= MODEL_TOL * k["grad_A_loss"]
        bad, tot = [], 0.0
        for key, p in m.named_parameters():
            if key not in k["grad_A"]:
                continue
            s, sq = k["grad_A"][key]
            gsq = float((p.grad.double() ** 2).sum())
            tot += gsq
            if "conv_block.conv" in key and key.endswith(".bias"):
                continue        # bias in front of an InstanceNorm: the true gradient is zero, both sides hold rounding noise
            if sq > 1e-16 and abs(gsq - sq) > GRAD_TOL * sq:
                bad.append((key, gsq, sq))
        assert not bad, bad[:5]
        assert abs(tot ** 0.5 - k["grad_A_total_norm"]) <= 1e-2 * k["grad_A_total_norm"]
    finally:
        N.set_conv_precision(None)
        torch.cuda.empty_cache()
