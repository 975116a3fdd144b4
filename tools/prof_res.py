"""UnetResBlock (conv3^3 -> IN -> LReLU -> conv3^3 -> IN -> + x -> LReLU) forward + backward at the decoder1 shape
(B=4, 160^3, 48 -> 48): per-phase timings with CUDA events; also usable under ncu."""
import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
B, R, C = 4, 160, 48
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator().manual_seed(0)
x = torch.randn(B, R, R, R, C, device='cuda').requires_grad_(True)
w1 = (torch.randn(C, C, 3, 3, 3, device='cuda') / (27 * C) ** 0.5).requires_grad_(True)
w2 = (torch.randn(C, C, 3, 3, 3, device='cuda') / (27 * C) ** 0.5).requires_grad_(True)
b1 = torch.zeros(C, device='cuda', requires_grad=True); b2 = torch.zeros(C, device='cuda', requires_grad=True)
dy = torch.randn(B, R, R, R, C, device='cuda')
for i in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = N.functional.ResBlockFn.apply(x, w1, b1, w2, b2, None, None, 0.01)
    e[1].record()
    y.backward(dy)
    e[2].record()
    torch.cuda.synchronize()
    print("ResBlock dec1: fwd %.2f ms, bwd %.2f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    x.grad = w1.grad = w2.grad = b1.grad = b2.grad = None
