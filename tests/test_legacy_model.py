"""The legacy `SwinTransformer_MAE3D` (reference swin_mae3d.py:417-1064; nerf-mae_b200/swin_mae3d_legacy.py) against golden vectors
recorded from the live reference (oracle/make_golden_legacy.py) and, on the GPU box, against the unmodified reference class itself
(baseline/_ref)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "golden_legacy.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _build(N, strategy="random"):
    torch.manual_seed(0)
    random.seed(0)
    return N.SwinTransformer_MAE3D([4, 4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [4, 4, 4], resolution=160, masking_prob=0.75,
                                   masking_strategy=strategy)


def _grid():
    g = torch.Generator().manual_seed(77)
    return torch.rand(4, 150, 160, 131, generator=g)


def test_legacy_state_dict_and_init(gold):
    import nerf_mae_b200 as N
    m = _build(N)
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(gold["keys"])
    for k in ("decoder_layers.0.weight", "decoder_layers.4.bias", "decoder_layers.12.weight", "mask_token", "stages.3.1.attn.qkv.weight"):
        s, a = gold["init." + k]
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)) and abs(float(sd[k].double().abs().sum()) - a) <= 1e-9 * a, k
    assert N.SwinTransformer_MAE3D is not N.SwinTransformer_MAE3D_New          # the legacy name is no longer an alias


@pytest.mark.parametrize("strategy", ["random", "grid", "block"])
def test_legacy_mask_bit_exact(gold, strategy):
    import nerf_mae_b200 as N
    random.seed(11)
    np.random.seed(12)
    m = N.draw_legacy_mask((40, 40, 40), 0.75, strategy)
    assert np.array_equal(np.packbits(m), gold[f"{strategy}.mask"])
    assert N.draw_legacy_mask((8, 8, 8), 0.75, None).sum() == 0                 # constructor default: the reference masks nothing


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16x3", "fp16"])
def test_legacy_encoder_decoder_golden(gold, mode):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import nerf_mae_b200 as N
    prev = N.set_conv_precision(mode)
    try:
        m = _build(N).cuda().eval()
        xb, ext = m.transform([_grid().cuda()])
        for strategy in ("random", "grid", "block"):
            m.sampling_strategy = strategy
            random.seed(11)
            np.random.seed(12)
            with torch.no_grad():
                latent, mask = m.forward_encoder(xb)
            assert np.array_equal(np.packbits(mask[0, ..., 0].cpu().numpy().astype(np.uint8)), gold[f"{strategy}.mask"])
            assert rel(latent.flatten()[torch.from_numpy(gold["idx_latent"]).cuda()], torch.from_numpy(gold[f"{strategy}.latent_sample"])) < 1e-3
            assert abs(float((latent.double() ** 2).sum()) - gold[f"{strategy}.latent_sums"][1]) <= 1e-3 * gold[f"{strategy}.latent_sums"][1]
            if strategy == "random":
                with torch.no_grad():
                    pred = m.forward_decoder(latent)
                assert list(pred.shape) == list(gold["random.pred_shape"])
                assert rel(pred.flatten()[torch.from_numpy(gold["idx_pred"]).cuda()], torch.from_numpy(gold["random.pred_sample"])) < 2e-3
                assert abs(float((pred.double() ** 2).sum()) - gold["random.pred_sums"][1]) <= 2e-3 * gold["random.pred_sums"][1]
        assert int(gold["forward_asserts"]) == 1
        with pytest.raises(AssertionError):                                         # exactly like the reference's forward()
            m([_grid().cuda()])
    finally:
        N.set_conv_precision(prev)


@pytest.mark.gpu
def test_legacy_decoder_gradients_vs_reference_on_gpu():
    """Decoder forward + backward against the unmodified reference class on the same GPU (strict fp32 math)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "nerf_mae", "model", "mae")):
        pytest.skip("baseline/_ref is not staged")
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from nerf_mae.model.mae import swin_mae3d as R
    import nerf_mae_b200 as N
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    prev = N.set_conv_precision("bf16x3")
    try:
        ours = _build(N).cuda().train()
        torch.manual_seed(0)
        random.seed(0)
        ref = R.SwinTransformer_MAE3D([4, 4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [4, 4, 4], resolution=160, masking_prob=0.75,
                                      masking_strategy="random").cuda().train()
        ref.load_state_dict(ours.state_dict())
        g = torch.Generator().manual_seed(3)
        latent = torch.randn(2, 5, 5, 5, 768, generator=g).cuda()
        la, lb = latent.clone().requires_grad_(True), latent.clone().requires_grad_(True)
        pa, pb = ours.forward_decoder(la), ref.forward_decoder(lb)
        assert rel(pa, pb) < 2e-4
        dy = torch.randn(pb.shape, generator=g).cuda()
        pa.backward(dy)
        pb.backward(dy)
        assert rel(la.grad, lb.grad) < 2e-2                                          # LeakyReLU(0.2) kinks: see test_gpu_parity.rel_trim
        for k in ("decoder_layers.0.weight", "decoder_layers.4.weight", "decoder_layers.8.weight", "decoder_layers.12.weight",
                  "decoder_layers.12.bias"):
            ga, gb = dict(ours.named_parameters())[k].grad, dict(ref.named_parameters())[k].grad
            assert rel(ga, gb) < 2e-2, k
    finally:
        N.set_conv_precision(prev)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
