"""Pins the oracle restatement against the LIVE reference (build container only: /root/reference does not travel)."""
import os
import random
import sys

import pytest
import torch

from oracle import nerf_mae_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nerf_mae")), reason="reference tree not present")


@pytest.fixture(scope="module")
def R():
    import numpy
    numpy.float = float          # torch_utils.py:42 needs the alias removed in numpy 2
    sys.path.insert(0, REF)
    from nerf_mae.model.mae import swin_mae3d
    return swin_mae3d


def test_forward_and_grads_match_live_reference(R):
    torch.manual_seed(3)
    m = R.SwinTransformer_MAE3D_New([4, 4, 4], 96, [2, 2, 2, 2], [3, 6, 12, 24], [4, 4, 4], resolution=32, masking_prob=0.75,
                                    stochastic_depth_prob=0.0).train()
    g = torch.Generator().manual_seed(5)
    grids = [torch.rand(4, 32, 30, 17, generator=g), torch.rand(4, 21, 32, 32, generator=g)]
    random.seed(9)
    loss, lr, la = m(grids)
    loss.backward()
    # the oracle is evaluated in float64: in fp32 its explicit InstanceNorm backward is ill-conditioned on
    # near-constant channels (3e-3 off), whereas the reference's fused native kernel stays at 2e-6 of the fp64 truth
    sd = {k: (v.detach().clone().double() if v.dtype.is_floating_point else v.clone()) for k, v in m.state_dict().items()}
    for k, v in sd.items():
        v.requires_grad_(v.dtype.is_floating_point and k != "pos_embed")
    torch.set_default_dtype(torch.float64)
    try:
        random.seed(9)
        lo, lro, lao = O.forward(sd, [x.double() for x in grids], [2, 2, 2, 2], [3, 6, 12, 24], 32, 0.75)
        lo.backward()
    finally:
        torch.set_default_dtype(torch.float32)
    for a, b in ((loss, lo), (lr, lro), (la, lao)):
        assert abs(float(a) - float(b)) <= 2e-6 * abs(float(a))
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        if "conv_block.conv" in k and k.endswith(".bias"):   # bias in front of InstanceNorm: exactly zero gradient
            assert float(p.grad.abs().max()) < 1e-4 and float(sd[k].grad.abs().max()) < 1e-9, k
            continue
        d = (p.grad.double() - sd[k].grad).norm() / (p.grad.norm() + 1e-12)
        assert float(d) < 5e-5 or float(p.grad.norm()) < 1e-6, (k, float(d))


def test_encoder_features_match_live_reference(R):
    """feature_extractor.py:1171-1184 (patch_partition + pos_embed + stages, no masking) vs oracle.encoder_features."""
    torch.manual_seed(4)
    m = R.SwinTransformer_MAE3D_New([4, 4, 4], 96, [2, 2, 2, 2], [3, 6, 12, 24], [4, 4, 4], resolution=32, masking_prob=0.75,
                                    stochastic_depth_prob=0.0).eval()
    x = torch.rand(2, 4, 32, 32, 32, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        t = m.patch_partition(x)
        t = t + m.pos_embed.type_as(t)
        ref = []
        for st in m.stages:
            t = st(t)
            ref.append(t.permute(0, 4, 1, 2, 3).contiguous())
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        got = O.encoder_features(sd, x, [2, 2, 2, 2], [3, 6, 12, 24])
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert float((a - b).norm() / b.norm()) < 2e-6


def test_scene_loading_and_augmentation_match_live_reference(R, tmp_path):
    """The driver fork's scene loader / augmentation (nerf-mae_b200/run_swin_mae3d.py) against nerf_rpn/datasets.py: same tensors,
    same number and order of Python RNG draws (float32 and uint8 scene files, every rotate/flip outcome)."""
    import numpy as np
    from nerf_rpn.datasets import BaseDataset
    import nerf_mae_b200  # noqa: F401  (registers the package)
    from nerf_mae_b200 import run_swin_mae3d as D
    rng = np.random.default_rng(0)
    scenes = {"f32": rng.normal(size=(9, 7, 5, 4)).astype(np.float32), "u8": rng.integers(0, 256, size=(6, 8, 4, 4), dtype=np.uint8)}
    for name, arr in scenes.items():
        np.savez(tmp_path / f"{name}.npz", rgbsigma=arr, resolution=np.asarray(arr.shape[:3]))
    ref_ds = BaseDataset(features_path=str(tmp_path), scene_list=list(scenes), normalize_density=True)
    for name in scenes:
        _, ref, _ = ref_ds.load_single_scene(name)
        got = D.load_scene_features(str(tmp_path / f"{name}.npz"), True)
        assert got.dtype == ref.dtype and torch.equal(got, ref), name
        for seed in range(12):
            random.seed(seed)
            a, _ = BaseDataset.augment_rpn_inputs(ref, None, 0.5, 0.5, 0.0, True)
            sa = random.random()
            random.seed(seed)
            b = D.augment_grid(got, 0.5, 0.5)
            sb = random.random()
            assert torch.equal(a, b) and sa == sb, (name, seed)


def test_checkpoints_interchange_with_live_reference(R, tmp_path):
    """A checkpoint written by the reference driver ({"epoch","state_dict","train_args"}, run_swin_mae3d.py:471-489) loads
    strictly into the drop-in model and the other way round: same keys, shapes and dtypes."""
    import nerf_mae_b200 as N
    cfg = dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=[4, 4, 4], resolution=32,
               masking_prob=0.75)
    torch.manual_seed(11)
    ref = R.SwinTransformer_MAE3D_New(**cfg)
    path = tmp_path / "ckpt.pt"
    torch.save({"epoch": 3, "state_dict": ref.state_dict(), "train_args": {"resolution": 32}}, path)
    ours = N.SwinTransformer_MAE3D_New(**cfg)
    ck = torch.load(path, map_location="cpu")
    missing, unexpected = ours.load_state_dict(ck["state_dict"], strict=True)
    assert not missing and not unexpected
    for k, v in ref.state_dict().items():
        w = ours.state_dict()[k]
        assert w.shape == v.shape and w.dtype == v.dtype and torch.equal(w, v), k
    # and back: a checkpoint of the drop-in model loads into the reference class
    torch.manual_seed(12)
    ours2 = N.SwinTransformer_MAE3D_New(**cfg)
    ref2 = R.SwinTransformer_MAE3D_New(**cfg)
    missing, unexpected = ref2.load_state_dict(ours2.state_dict(), strict=True)
    assert not missing and not unexpected
