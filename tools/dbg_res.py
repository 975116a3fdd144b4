import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from oracle import nerf_mae_oracle as O
torch.manual_seed(0)
g = torch.Generator().manual_seed(21)
blk = N.UnetResBlock(48, 48, 3).cuda()
sd = {k: v.detach().cpu().double() for k, v in blk.state_dict().items()}
xin = torch.randn(2, 48, 24, 20, 28, generator=g)
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())
# oracle pieces in float64
torch.set_default_dtype(torch.float64)
xo = xin.double().requires_grad_(True)
import torch.nn.functional as F
y1 = F.conv3d(xo, sd['conv1.weight'], sd['conv1.bias'], padding=1); y1.retain_grad()
a1 = O.leaky_relu(O.instance_norm_cl(y1)); a1.retain_grad()
y2 = F.conv3d(a1, sd['conv2.weight'], sd['conv2.bias'], padding=1); y2.retain_grad()
out = O.leaky_relu(O.instance_norm_cl(y2) + xo)
dy = torch.randn(out.shape, generator=g, dtype=torch.float32).double()
out.backward(dy)
torch.set_default_dtype(torch.float32)
# ours, step by step through the C ABI
from nerf_mae_b200._lib import call
cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
x = cl(xin).cuda(); B, X, Y, Z, C = x.shape; V = X * Y * Z; dev = x.device
w1, b1, w2, b2 = blk.conv1.weight.data, blk.conv1.bias.data, blk.conv2.weight.data, blk.conv2.bias.data
wws = torch.empty(27 * C * C, device=dev)
Y1 = torch.empty_like(x); st1 = torch.empty(B, C, 2, dtype=torch.float64, device=dev)
call("nmae_conv3x3x3_fwd", x, w1, b1, B, X, Y, Z, C, C, wws, Y1, device=dev)
call("nmae_instnorm_stats", Y1, B, V, C, st1, device=dev)
A1 = torch.empty_like(x); call("nmae_in_lrelu_apply_fwd", Y1, st1, None, None, B, V, C, 1e-5, 0.01, A1, device=dev)
Y2 = torch.empty_like(x); st2 = torch.empty_like(st1)
call("nmae_conv3x3x3_fwd", A1, w2, b2, B, X, Y, Z, C, C, wws, Y2, device=dev)
call("nmae_instnorm_stats", Y2, B, V, C, st2, device=dev)
OUT = torch.empty_like(x); call("nmae_in_lrelu_apply_fwd", Y2, st2, x, None, B, V, C, 1e-5, 0.01, OUT, device=dev)
print('fwd y1', rel(Y1, cl(y1)), 'a1', rel(A1, cl(a1)), 'y2', rel(Y2, cl(y2)), 'out', rel(OUT, cl(out)))
DOUT = cl(dy.float()).cuda()
sums = torch.empty(B, C, 3, dtype=torch.float64, device=dev)
DY2 = torch.empty_like(x); DX = torch.empty_like(x)
call("nmae_in_lrelu_apply_bwd", DOUT, OUT, Y2, st2, None, None, B, V, C, 1e-5, 0.01, sums, DY2, None, DX, device=dev)
print('dy2', rel(DY2, cl(y2.grad)))
DA1 = torch.empty_like(x)
call("nmae_conv3x3x3_dgrad", DY2, w2, B, X, Y, Z, C, C, wws, DA1, 0, device=dev)
print('da1', rel(DA1, cl(a1.grad)))
# same with the exact dy2
DA1b = torch.empty_like(x)
call("nmae_conv3x3x3_dgrad", cl(y2.grad.float()).cuda(), w2, B, X, Y, Z, C, C, wws, DA1b, 0, device=dev)
print('da1 (exact input)', rel(DA1b, cl(a1.grad)))
DY1 = torch.empty_like(x)
call("nmae_in_lrelu_apply_bwd", DA1, A1, Y1, st1, None, None, B, V, C, 1e-5, 0.01, sums, DY1, None, None, device=dev)
print('dy1', rel(DY1, cl(y1.grad)))
DY1b = torch.empty_like(x)
call("nmae_in_lrelu_apply_bwd", cl(a1.grad.float()).cuda(), cl(a1.float()).cuda(), cl(y1.float()).cuda(), st1, None, None, B, V, C, 1e-5, 0.01, sums, DY1b, None, None, device=dev)
print('dy1 (exact inputs)', rel(DY1b, cl(y1.grad)))
call("nmae_conv3x3x3_dgrad", DY1, w1, B, X, Y, Z, C, C, wws, DX, 1, device=dev)
print('dx', rel(DX, cl(xo.grad)))
print('y2.grad norm', float(y2.grad.norm()), 'a1.grad norm', float(a1.grad.norm()), 'y1.grad', float(y1.grad.norm()))
