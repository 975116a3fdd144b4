"""The CPU oracle against fixtures generated from the live reference (oracle/make_golden.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import nerf_mae_oracle as O

T = torch.from_numpy


@pytest.mark.parametrize("name", ["plain", "shift", "pad_shift", "pad5_shift", "ragged_shift", "tiny_noshift"])
def test_window_attention(golden, name):
    g = {k.split(".")[-1]: v for k, v in golden.items() if k.startswith(f"attn.{name}.")}
    nh, sh = (int(v) for v in g["meta"])
    y = O.window_attention(T(g["x"]), T(g["qw"]), T(g["qb"]), T(g["pw"]), T(g["pb"]), T(g["table"]), nh, 4, sh)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=2e-5)


def test_rel_index(golden):
    assert np.array_equal(O.relative_position_index(4).numpy(), golden["attn.rel_index"])  # bit-exact ints


@pytest.mark.parametrize("name", ["even", "odd", "ragged"])
def test_patch_merge(golden, name):
    g = {k.split(".")[-1]: T(v) for k, v in golden.items() if k.startswith(f"merge.{name}.")}
    sd = {"norm.weight": g["nw"], "norm.bias": g["nb"], "reduction.weight": g["rw"]}
    np.testing.assert_allclose(O.patch_merge(g["x"], sd, "").numpy(), g["y"].numpy(), rtol=1e-4, atol=1e-5)


def test_swin_block(golden):
    sd = {k[len("block.sd."):]: T(v) for k, v in golden.items() if k.startswith("block.sd.")}
    y = O.swin_block(T(golden["block.x"]), sd, "", 1, 2)
    np.testing.assert_allclose(y.numpy(), golden["block.y"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["skip", "noskip"])
def test_up_block(golden, name):
    sd = {k[len(f"up.{name}.sd."):]: T(v) for k, v in golden.items() if k.startswith(f"up.{name}.sd.")}
    skip = T(golden["up.skip.skip"]) if name == "skip" else None
    y = O.up_block(T(golden[f"up.{name}.x"]), skip, sd, "")
    np.testing.assert_allclose(y.numpy(), golden[f"up.{name}.y"], rtol=1e-4, atol=1e-5)


def test_pos_embed_and_pad(golden):
    np.testing.assert_allclose(O.sincos_pos_embed_3d(96, 5).numpy(), golden["pos_embed.96.5"], atol=1e-6)
    np.testing.assert_allclose(O.sincos_pos_embed_3d(192, 3).numpy(), golden["pos_embed.192.3"], atol=1e-6)
    x, ext = O.pad_grids([T(golden["pad.in"])], 8)
    assert np.array_equal(x.numpy(), golden["pad.out"])
    m = golden["pad.mask"]
    assert ext.tolist() == [[3, 5, 2]] and m[0, :, :3, :5, :2].all() and m.sum() == 4 * 3 * 5 * 2


@pytest.mark.parametrize("n_tok,seed", [(40, 123), (16, 42), (10, 3)])
def test_mask_bit_exact(golden, n_tok, seed):
    random.seed(seed)
    m = O.draw_block_mask((n_tok,) * 3, 0.75)
    assert np.array_equal(np.packbits(m.numpy().astype(np.uint8)), golden[f"mask.{n_tok}.{seed}"])


def test_mask_empty_and_full():
    assert not O.draw_block_mask((3, 3, 3), 0.75).any()            # smaller than one block: never masked
    assert O.draw_block_mask((8, 8, 8), 1.1).all() and not O.draw_block_mask((8, 8, 8), 0.0).any()


@pytest.mark.parametrize("tag", ["even", "odd"])
def test_fpn(golden_fpn, tag):
    """fpn.py:134-185 restated (oracle.fpn_forward) against the live reference's outputs: 2x pyramid and non-2x sizes."""
    sd = {k[len("fpn.sd."):]: T(v) for k, v in golden_fpn.items() if k.startswith("fpn.sd.")}
    feats = [T(golden_fpn[f"fpn.{tag}.x{i}"]) for i in range(4)]
    for i, y in enumerate(O.fpn_forward(sd, feats)):
        np.testing.assert_allclose(y.numpy(), golden_fpn[f"fpn.{tag}.y{i}"], rtol=1e-4, atol=1e-5)
