"""Runs the W-MSA core (tcgen05) forward + backward at the stage-1 and stage-3 shapes of swin_s @160^3, B=4; used under ncu."""
import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200._lib import call, num_windows
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (B, H, C, nH, shift) in [(4, 40, 96, 3, 2), (4, 10, 384, 12, 2)]:
    T = H * H * H; M = B * T
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(M + 1, 3 * C, generator=g).cuda()
    table = (torch.randn(343, nH, generator=g) * 0.02).cuda()
    dout = torch.randn(M, C, generator=g).cuda()
    nW = num_windows(H, H, H)
    out = torch.empty(M, C, device='cuda'); lse = torch.empty(B * nW * nH * 64, device='cuda')
    dqkv = torch.empty_like(qkv); dtable = torch.empty_like(table)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(reps):
        ev[0].record()
        call("nmae_window_attention_fwd", qkv, table, B, H, H, H, C, nH, shift, out, lse, device=qkv.device)
        ev[1].record()
        call("nmae_window_attention_bwd", dout, qkv, table, out, lse, B, H, H, H, C, nH, shift, dqkv, dtable, device=qkv.device)
        ev[2].record()
    torch.cuda.synchronize()
    flop_f = 2 * 2 * 64 * 64 * 32 * B * nW * nH
    print("B=%d H=%d C=%d heads=%d: fwd %.3f ms bwd %.3f ms  (useful %.1f / %.1f TFLOP/s)" % (
        B, H, C, nH, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), flop_f / ev[0].elapsed_time(ev[1]) / 1e9,
        2.5 * flop_f / ev[1].elapsed_time(ev[2]) / 1e9))
