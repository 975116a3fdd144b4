"""Per-entry-point GPU time of one training step (swin_s, 4 x 160^3): every C-ABI call is bracketed with CUDA events.
    python tools/step_breakdown.py [steps=3]     (NMAE_USE_DBG_LIB=1 selects the -DNMAE_DBG build for NMAE_DBG experiments)"""
import collections
import os
import random
import sys

import torch

sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200 import _lib
from nerf_mae_b200.trainer import MAEStepper

if os.environ.get("NMAE_NO_FUSE_OUT"):
    from nerf_mae_b200 import unetr_block
    unetr_block.UnetResBlock.can_fuse_out = lambda self, x, o: False
if os.environ.get("NMAE_USE_DBG_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libnmae_dbg.so")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model_name = os.environ.get("PM", "swin_s")
torch.manual_seed(0)
model = N.build_model(model_name, 160, 0.75).cuda().train()
st = MAEStepper(model, total_steps=64)
gen = torch.Generator().manual_seed(0)
grids = [torch.rand(4, 160, 160, 160, generator=gen).cuda() for _ in range(4)]
random.seed(0)
for _ in range(2):
    st.step(grids)
torch.cuda.synchronize()


class Rec(dict):
    def __contains__(self, k):
        return True

    def __missing__(self, k):
        self[k] = []
        return self[k]


_lib.timed_calls = Rec()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    st.step(grids)
e1.record()
torch.cuda.synchronize()
calls, _lib.timed_calls = _lib.timed_calls, None
tot = e0.elapsed_time(e1) / steps
agg = collections.defaultdict(lambda: [0, 0.0])
big = collections.defaultdict(lambda: [0, 0.0])
for name, evs in calls.items():
    for a, b, ints in evs:
        t = a.elapsed_time(b)
        agg[name][0] += 1
        agg[name][1] += t
        if name in ("nmae_linear_fwd", "nmae_linear_bwd_input", "nmae_linear_bwd_weight", "nmae_in_lrelu_apply_bwd_image_h",
                    "nmae_convT_k_eq_s_fwd", "nmae_convT_k_eq_s_bwd", "nmae_instnorm_stats", "nmae_in_lrelu_apply_fwd"):
            key = (name, ints[:5] if "linear" not in name else ints[:3])
            big[key][0] += 1
            big[key][1] += t
print("NMAE_DBG=%s  %.2f ms/step (with per-call events)" % (os.environ.get("NMAE_DBG", "0"), tot))
s = 0.0
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-34s %5d calls/step %8.3f ms/step" % (name, n // steps, t / steps))
    s += t / steps
print("  sum %.2f" % s)
print("linear shapes (M,N,K):")
for (name, shp), (n, t) in sorted(big.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('PTOP', 16))]:
    print("  %-24s %-22s %3d/step %7.3f ms/step" % (name, shp, n // steps, t / steps))
