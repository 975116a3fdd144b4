"""Runs the decoder1-geometry 3x3x3 convolution kernels (B=4, 160^3, 48->48) a few times; used under ncu."""
import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200._lib import call
import os
B, R, C = int(os.environ.get('PB', 4)), 160, 48
g = torch.Generator().manual_seed(0)
x = torch.randn(B, R, R, R, C, device='cuda')
dy = torch.randn(B, R, R, R, C, device='cuda')
w = torch.randn(C, C, 3, 3, 3, device='cuda') / (27 * C) ** 0.5
b = torch.randn(C, device='cuda')
wws = torch.empty(27 * C * C, device='cuda')
y = torch.empty_like(x); dw = torch.empty_like(w); db = torch.empty_like(b)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(reps):
    ev[0].record()
    ximg = N.functional.conv3_image(x)
    evi = torch.cuda.Event(enable_timing=True); evi.record()
    call("nmae_conv3x3x3_fwd", x, ximg, w, b, B, R, R, R, C, C, wws, y, device=x.device)
    ev[1].record()
    dyimg = N.functional.conv3_image(dy)
    evd = torch.cuda.Event(enable_timing=True); evd.record()
    call("nmae_conv3x3x3_dgrad", dy, dyimg, w, B, R, R, R, C, C, wws, y, 0, device=x.device)
    ev[2].record()
    call("nmae_conv3x3x3_wgrad", dy, dyimg, x, ximg, B, R, R, R, C, C, wws, dw, db, device=x.device)
    ev[3].record()
torch.cuda.synchronize()
print("ms image %.2f fwd %.2f dgrad %.2f wgrad %.2f" % (ev[0].elapsed_time(evi), evi.elapsed_time(ev[1]), evd.elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
