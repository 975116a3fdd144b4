"""Boundary tests on the GPU: the claims of INTEGRATION.md that no per-operator test covers.

  * section 2: this repo's blocks injected into the UNMODIFIED reference model class through its constructor callables
    (`block=`, `downsample_layer=`, `norm_layer=`, swin_mae3d.py:1099-1103) - needs baseline/_ref (staged by build());
  * section 1: `DistributedDataParallel(model)` around the drop-in model trains (single-process NCCL group);
  * the driver fork (nerf-mae_b200/run_swin_mae3d.py): synthetic train -> checkpoint -> `--checkpoint` resume reproduces the
    next epoch's losses (optimizer moments + step count, OneCycleLR position, Python / torch / CUDA RNG streams) -> eval.
"""
import os
import random
import sys
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def N():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import nerf_mae_b200
    nerf_mae_b200.lib()
    return nerf_mae_b200


def _reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "nerf_mae", "model", "mae")):
        pytest.skip("baseline/_ref is not staged (python -c 'import __graft_entry__ as g; g.build()' in the build container)")
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from nerf_mae.model.mae import swin_mae3d as R
    return R


def test_blocks_injected_into_reference_model(N):
    R = _reference()
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False      # strict fp32 reference
    try:
        kw = dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=[4, 4, 4],
                  resolution=64, masking_prob=0.75)
        torch.manual_seed(0)
        ref = R.SwinTransformer_MAE3D_New(**kw).cuda().eval()
        torch.manual_seed(0)
        inj = R.SwinTransformer_MAE3D_New(**kw, norm_layer=partial(N.LayerNorm, eps=1e-5),
                                          block=partial(N.SwinTransformerBlock, attn_layer=N.ShiftedWindowAttention),
                                          downsample_layer=N.PatchMerging).cuda().eval()
        assert set(inj.state_dict()) == set(ref.state_dict())
        inj.load_state_dict(ref.state_dict())
        g = torch.Generator().manual_seed(4)
        grids = [torch.rand(4, 64, 64, 64, generator=g).cuda(), torch.rand(4, 50, 64, 41, generator=g).cuda()]
        with torch.no_grad():
            random.seed(7)
            a = ref(grids, is_eval=True)
            random.seed(7)
            b = inj(grids, is_eval=True)
        for x, y in zip(a[:3], b[:3]):
            assert abs(float(x) - float(y)) <= 1e-3 * abs(float(x))
        assert float((a[3] - b[3]).norm() / a[3].norm()) < 1e-3
        # and it trains: gradients flow through the injected blocks into the reference's parameters
        inj.train()
        random.seed(7)
        loss, _, _ = inj(grids)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in inj.parameters() if p.requires_grad)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


def test_ddp_wrapped_model_steps(N):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29631", rank=0, world_size=1)
    try:
        torch.manual_seed(0)
        m = N.build_model("swin_t", 32, 0.75, stochastic_depth_prob=0.0).cuda().train()
        g = torch.Generator().manual_seed(1)
        grids = [torch.rand(4, 32, 32, 32, generator=g).cuda(), torch.rand(4, 20, 32, 27, generator=g).cuda()]
        random.seed(3)
        want, _, _ = m(grids)
        want.backward()
        gref = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        m.zero_grad(set_to_none=True)
        ddp = DDP(m, device_ids=[torch.cuda.current_device()])
        opt = N.FusedAdamWClip([p for p in ddp.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3, clip_grad_norm=0.1)
        random.seed(3)
        loss, _, _ = ddp(grids)
        loss.backward()
        assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
        for k, p in m.named_parameters():
            if "conv_block.conv" in k and k.endswith(".bias"):
                continue        # bias in front of an InstanceNorm: exactly-zero gradient, both runs hold rounding noise
            if k in gref:
                assert float((p.grad - gref[k]).norm()) <= 1e-3 * float(gref[k].norm()) + 1e-12, k
        opt.step()
        assert torch.isfinite(opt.grad_norm()).item()
    finally:
        if own_group:
            dist.destroy_process_group()


def test_driver_train_resume_eval(N, tmp_path):
    from nerf_mae_b200 import run_swin_mae3d as D

    def args(save, extra=()):
        return D.parse_args(["--mode", "train", "--dataset", "synthetic", "--synthetic_scenes", "5", "--resolution", "32",
                             "--backbone_type", "swin_t", "--batch_size", "2", "--num_epochs", "2", "--lr", "1e-4",
                             "--weight_decay", "1e-3", "--masking_prob", "0.75", "--log_interval", "1", "--eval_interval", "1",
                             "--keep_checkpoints", "2", "--save_path", str(save)] + list(extra))

    # run A: two epochs in one go
    full = D.Trainer(args(tmp_path / "a"), 0, 1, torch.cuda.current_device())
    full.train_loop()
    hist_full = [h for h in full.history if h[0] == 1]
    assert len(hist_full) >= 2 and os.path.exists(tmp_path / "a" / "epoch_0.pt") and os.path.exists(tmp_path / "a" / "epoch_1.pt")
    # run B: resume from the epoch-0 checkpoint of run A and train epoch 1 only
    res = D.Trainer(args(tmp_path / "b", ["--checkpoint", str(tmp_path / "a" / "epoch_0.pt")]), 0, 1, torch.cuda.current_device())
    assert res.start_epoch == 1
    res.train_loop()
    hist_res = [h for h in res.history if h[0] == 1]
    assert len(hist_res) == len(hist_full)
    for a, b in zip(hist_full, hist_res):
        assert abs(a[2] - b[2]) <= 2e-4 * abs(a[2]), (a, b)          # same data order, masks, stochastic depth, lr, moments
    assert res.optimizer._steps == full.optimizer._steps
    # eval mode on the saved checkpoint
    ev = D.Trainer(D.parse_args(["--mode", "eval", "--dataset", "synthetic", "--synthetic_scenes", "5", "--resolution", "32",
                                 "--backbone_type", "swin_t", "--batch_size", "2", "--masking_prob", "0.75",
                                 "--checkpoint", str(tmp_path / "b" / "epoch_1.pt"), "--save_path", str(tmp_path / "e")]),
                   0, 1, torch.cuda.current_device())
    _, val = D.scene_lists(ev.args)
    out = ev.eval(D.SceneDataset(ev.args, val, False))
    assert set(out) == {"psnr", "mse", "loss"} and all(torch.isfinite(torch.tensor(v)) for v in out.values())
