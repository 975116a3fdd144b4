"""nerf-mae_b200: B200-native (sm_100a) 3D Swin-MAE pretraining hot path behind the reference's module API.

The directory name is not a Python identifier; import it as `nerf_mae_b200` through the shim module
`nerf_mae_b200.py` at the repository root.
"""
from . import functional  # noqa: F401
from .functional import get_conv_precision, set_conv_precision  # noqa: F401
from ._lib import LIB_PATH, exported_symbols, lib  # noqa: F401
from .optim import FusedAdamWClip, GradAllReducer  # noqa: F401
from .swin_mae3d import (SWIN_CONFIGS, LayerNorm, PatchMerging, ShiftedWindowAttention,  # noqa: F401
                         SwinTransformer_MAE3D_New, SwinTransformerBlock, build_model, draw_block_mask,
                         shifted_window_attention)
from .swin_mae3d_legacy import SwinTransformer_MAE3D, draw_legacy_mask  # noqa: F401
from .unetr_block import UnetOutBlock, UnetResBlock, UnetrUpBlock  # noqa: F401
from .fpn import FPN, SwinTransformer_FPN_Pretrained, SwinTransformer_FPN_Pretrained_Skip  # noqa: F401

__version__ = "0.1.0"
