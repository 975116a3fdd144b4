"""Runs the tcgen05 linear kernels at representative shapes of swin_s @160^3, B=4 (stage-1 qkv / fc1, stage-3 qkv / fc2, the
decoder1 transposed convolution); used under ncu and for quick timings."""
import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200 import _lib
import os
if os.environ.get("NMAE_USE_DBG_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libnmae_dbg.so")
if os.environ.get("NMAE_LIB_PATH"):     # same-box A/B against another build of the library
    _lib.LIB_PATH = os.environ["NMAE_LIB_PATH"]
call = _lib.call
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator().manual_seed(0)
for (M, K, Nn, tag) in [(256000, 96, 288, "stage1 qkv"), (256000, 96, 384, "stage1 fc1"), (256000, 384, 96, "stage1 fc2"),
                        (4000, 384, 1152, "stage3 qkv"), (4000, 384, 1536, "stage3 fc1"), (4000, 1536, 384, "stage3 fc2")]:
    x = torch.randn(M, K, generator=g).cuda(); w = (torch.randn(Nn, K, generator=g) * 0.02).cuda(); b = torch.zeros(Nn).cuda()
    dy = torch.randn(M, Nn, generator=g).cuda()
    y = torch.empty(M, Nn, device='cuda'); dx = torch.empty_like(x); dw = torch.empty_like(w); db = torch.empty_like(b)
    ws = torch.empty(w.numel(), device='cuda')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for i in range(reps):
        ev[0].record()
        call("nmae_linear_fwd", x, w, b, M, Nn, K, 0, None, None, None, 1, y, ws, device=x.device)
        ev[1].record()
        call("nmae_linear_bwd_input", dy, w, M, Nn, K, 0, None, dx, ws, device=x.device)
        ev[2].record()
        call("nmae_linear_bwd_weight", dy, x, M, Nn, K, dw, db, device=x.device)
        ev[3].record()
    torch.cuda.synchronize()
    fl = 2.0 * M * K * Nn
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    print("%-11s M=%6d K=%4d N=%4d: fwd %.3f ms (%.0f TF/s) dgrad %.3f ms (%.0f) wgrad+bias %.3f ms (%.0f); fwd bytes %.0f MB -> %.0f GB/s" % (
        tag, M, K, Nn, t[0], fl / t[0] / 1e9, t[1], fl / t[1] / 1e9, t[2], fl / t[2] / 1e9, (M * K + M * Nn) * 4 / 1e6, (M * K + M * Nn) * 4 / t[0] / 1e6))

# the GELU epilogues (fc1 forward saves the pre-activation; fc2's input gradient multiplies by gelu'(pre-activation))
for (M, C, tag) in [(256000, 96, "stage1"), (32000, 192, "stage2"), (4000, 384, "stage3")]:
    H = 4 * C
    x = torch.randn(M, C, generator=g).cuda(); w1 = (torch.randn(H, C, generator=g) * 0.02).cuda(); b1 = torch.zeros(H).cuda()
    w2 = (torch.randn(C, H, generator=g) * 0.02).cuda(); dy = torch.randn(M, C, generator=g).cuda()
    aux = torch.empty(M, H, device='cuda'); h = torch.empty(M, H, device='cuda'); dh = torch.empty(M, H, device='cuda')
    ws1 = torch.empty(w1.numel(), device='cuda'); ws2 = torch.empty(w2.numel(), device='cuda')
    b2 = torch.zeros(C).cuda(); y = torch.empty(M, C, device='cuda'); dx = torch.zeros(M, C, device='cuda')
    rs = torch.ones(4, device='cuda')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for i in range(reps):
        ev[0].record()
        call("nmae_linear_fwd", x, w1, b1, M, H, C, 1, aux, None, None, 1, h, ws1, device=x.device)
        ev[1].record()
        call("nmae_linear_bwd_input", dy, w2, M, C, H, 1, aux, dh, ws2, device=x.device)
        ev[2].record()
        call("nmae_linear_fwd", h, w2, b2, M, C, H, 2, None, x, rs, M // 4, y, ws2, device=x.device)
        ev[3].record()
        call("nmae_linear_bwd_input", dh, w1, M, H, C, 4, None, dx, ws1, device=x.device)
        ev[4].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    print("%-7s MLP M=%6d C=%4d: fc1 fwd + GELU %.3f ms, fc2 dgrad * GELU' %.3f ms, fc2 fwd + residual %.3f ms, fc1 dgrad accumulate %.3f ms" % (
        tag, M, C, t[0], t[1], t[2], t[3]))
