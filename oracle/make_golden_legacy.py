"""Golden vectors for the LEGACY model `SwinTransformer_MAE3D` (nerf_mae/model/mae/swin_mae3d.py:417-1064) from the live reference.
Test infrastructure; build container only:  python oracle/make_golden_legacy.py  -> tests/golden/golden_legacy.npz

Recorded (swin_t, resolution 160 - the decoder's upsampling sizes are hard-coded for it, CPU fp32): for each masking strategy
("random", "grid", "block") the packed token mask, and for the "random" one the encoder latent and decoder output (checksums and
4096 sampled values); the state-dict key list and the init fingerprint of the decoder layers; the fact that forward() asserts.
"""
import os
import random
import sys

import numpy
import numpy as np
import torch

numpy.float = float
sys.path.insert(0, "/root/reference")
from nerf_mae.model.mae import swin_mae3d as R  # noqa: E402
from nerf_mae.model.mae.torch_utils import pad_tensor  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "golden_legacy.npz")

if __name__ == "__main__":
    a = {}
    torch.manual_seed(0)
    random.seed(0)
    m = R.SwinTransformer_MAE3D([4, 4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [4, 4, 4], resolution=160, masking_prob=0.75,
                                masking_strategy="random").eval()
    sd = m.state_dict()
    a["keys"] = np.array(sorted(sd.keys()))
    for k in ("decoder_layers.0.weight", "decoder_layers.4.bias", "decoder_layers.12.weight", "mask_token", "stages.3.1.attn.qkv.weight"):
        a["init." + k] = np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum())])
    g = torch.Generator().manual_seed(77)
    x = torch.rand(4, 150, 160, 131, generator=g)
    xb, _ = pad_tensor(x, [160, 160, 160], 0)
    idx_l = torch.randint(0, 5 ** 3 * 768, (4096,), generator=torch.Generator().manual_seed(5))
    idx_p = torch.randint(0, 40 ** 3 * 256, (4096,), generator=torch.Generator().manual_seed(6))
    a["idx_latent"], a["idx_pred"] = idx_l.numpy(), idx_p.numpy()
    for strat in ("random", "grid", "block"):
        m.sampling_strategy = strat
        random.seed(11)
        np.random.seed(12)
        with torch.no_grad():
            latent, mask = m.forward_encoder(xb)
            a[f"{strat}.mask"] = np.packbits(mask[0, ..., 0].numpy().astype(np.uint8))
            a[f"{strat}.latent_sample"] = latent.flatten()[idx_l].numpy()
            a[f"{strat}.latent_sums"] = np.array([float(latent.double().sum()), float((latent.double() ** 2).sum())])
            if strat == "random":
                pred = m.forward_decoder(latent)
                a["random.pred_shape"] = np.array(pred.shape)
                a["random.pred_sample"] = pred.flatten()[idx_p].numpy()
                a["random.pred_sums"] = np.array([float(pred.double().sum()), float((pred.double() ** 2).sum())])
        print(strat, int(mask.sum()), float(latent.abs().mean()), flush=True)
    try:
        m([x])
        a["forward_asserts"] = np.array(0)
    except AssertionError:
        a["forward_asserts"] = np.array(1)
    np.savez_compressed(OUT, **a)
    print("wrote", OUT, os.path.getsize(OUT))
