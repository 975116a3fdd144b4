"""Times the InstanceNorm+LeakyReLU backward (sums + apply-to-image) and the other HBM-bound decoder kernels at the decoder1
geometry (B=4, 160^3, 48 channels) in isolation; used under ncu as well."""
import os
import sys

import torch

sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200 import _lib
from nerf_mae_b200._lib import call, conv3h_image_bytes

B, R, C = 4, 160, 48
V = R ** 3
dev = torch.device("cuda")
dout = torch.randn(B, R, R, R, C, device=dev) * 1e-6
out = torch.randn(B, R, R, R, C, device=dev)
y2 = torch.randn(B, R, R, R, C, device=dev)
dp4 = torch.randn(B, R, R, R, 4, device=dev) * 1e-6
w_out = torch.randn(4, C, device=dev)
st = torch.empty(B, C, 2, dtype=torch.float64, device=dev)
sums = torch.empty(_lib.workspace_bytes('nmae_in_lrelu_bwd_sums_ws_bytes', B, C) // 8, dtype=torch.float64, device=dev)
scal = torch.empty(4, device=dev)
dres = torch.empty_like(out)
db = torch.empty(C, device=dev)
dwo, dbo = torch.empty(4, C, device=dev), torch.empty(4, device=dev)
img = torch.empty(conv3h_image_bytes(B, R, R, R, C), dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def timed(label, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %.3f ms" % (label, e0.elapsed_time(e1) / reps))


timed("instnorm_stats", lambda: call("nmae_instnorm_stats", y2, B, V, C, st, device=dev))
timed("in_lrelu_apply_fwd (+identity residual)", lambda: call("nmae_in_lrelu_apply_fwd", y2, st, out, None, B, V, C, 1e-5, 0.01, dres, device=dev))
pred = torch.empty(B, R, R, R, 4, device=dev)
bo = torch.zeros(4, device=dev)
timed("in_lrelu_apply_out_fwd (+identity residual, out conv)", lambda: call("nmae_in_lrelu_apply_out_fwd", y2, st, out, None, B, V, C, 1e-5, 0.01, dres,
                                                                            w_out, bo, pred, device=dev))
timed("out conv alone (thin forward)", lambda: call("nmae_linear_fwd", dres, w_out, bo, B * V, 4, C, 0, None, None, None, 1, pred, None, device=dev))
timed("conv3h_image_build (IN+LReLU fused)", lambda: call("nmae_conv3h_image_build", y2, C, 0, B, R, R, R, C, st, 1e-5, 0.01, None, img, device=dev))
timed("in_bwd_image_h conv2 (dout, out, dres)", lambda: call("nmae_in_lrelu_apply_bwd_image_h", dout, out, y2, st, None, None, B, R, R, R, C, 1e-5,
                                                              0.01, sums, scal[0:1], img, scal[1:2], None, dres, db, None, None, None, None, None, device=dev))
timed("in_bwd_image_h conv2 fused out-conv (dp4)", lambda: call("nmae_in_lrelu_apply_bwd_image_h", None, out, y2, st, None, None, B, R, R, R, C,
                                                                 1e-5, 0.01, sums, scal[0:1], img, scal[1:2], None, dres, db, None, dp4, w_out, dwo, dbo, device=dev))
timed("in_bwd_image_h conv1 (dout only)", lambda: call("nmae_in_lrelu_apply_bwd_image_h", dout, None, y2, st, None, None, B, R, R, R, C, 1e-5,
                                                        0.01, sums, scal[0:1], img, scal[1:2], None, None, db, None, None, None, None, None, device=dev))
