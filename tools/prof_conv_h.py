"""Times the fp16 single-pass convolution kernels at the decoder1 geometry (B=4, 160^3, 48->48); used for NMAE_DBG bottleneck
experiments (needs the -DNMAE_DBG build: python -c 'import __graft_entry__ as g; g.build_debug()') and under ncu."""
import os
import sys

import torch

sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200 import _lib

if os.environ.get("NMAE_USE_DBG_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libnmae_dbg.so")
call = _lib.call
B, R, C = int(os.environ.get('PB', 4)), int(os.environ.get('PR', 160)), int(os.environ.get('PC', 48))
x = torch.randn(B, R, R, R, C, device='cuda')
dy = torch.randn(B, R, R, R, C, device='cuda')
w = torch.randn(C, C, 3, 3, 3, device='cuda') / (27 * C) ** 0.5
b = torch.randn(C, device='cuda')
wws = torch.empty(27 * C * C, device='cuda')
y = torch.empty_like(x); dw = torch.empty_like(w)
one = torch.ones(1, device='cuda')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for i in range(reps):
    ev[0].record()
    ximg = N.functional.conv3h_image(x)
    ev[1].record()
    call("nmae_conv3h_fwd", ximg, w, b, B, R, R, R, C, C, wws, y, device=x.device)
    ev[2].record()
    dyimg = N.functional.conv3h_image(dy)
    ev[3].record()
    call("nmae_conv3h_dgrad", dyimg, one, w, B, R, R, R, C, C, wws, y, 0, device=x.device)
    ev[4].record()
    e5 = torch.cuda.Event(enable_timing=True)
    call("nmae_conv3h_wgrad", dyimg, one, ximg, B, R, R, R, C, C, dw, device=x.device)
    e5.record()
torch.cuda.synchronize()
print("NMAE_DBG=%s ms image %.2f fwd %.2f dgrad %.2f wgrad %.2f" % (os.environ.get("NMAE_DBG", "0"), ev[0].elapsed_time(ev[1]),
      ev[1].elapsed_time(ev[2]), ev[3].elapsed_time(ev[4]), ev[4].elapsed_time(e5)))
