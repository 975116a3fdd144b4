import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200._lib import call
def run(B, Ci, Co, X, Y, Z):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, X, Y, Z, Ci, generator=g).cuda()
    dy = torch.randn(B, X, Y, Z, Co, generator=g).cuda()
    wws = torch.empty(27 * Ci * Co, device='cuda')
    dw = torch.empty(Co, Ci, 3, 3, 3, device='cuda'); db = torch.empty(Co, device='cuda')
    call("nmae_conv3x3x3_wgrad", dy, x, B, X, Y, Z, Ci, Co, wws, dw, db, device=x.device)
    xr = x.permute(0, 4, 1, 2, 3).double().cpu(); dyr = dy.permute(0, 4, 1, 2, 3).double().cpu()
    ref = torch.nn.grad.conv3d_weight(xr, (Co, Ci, 3, 3, 3), dyr, padding=1)
    e = float((dw.double().cpu() - ref).norm() / ref.norm())
    per_tap = ((dw.double().cpu() - ref) ** 2).sum(dim=(0, 1)).sqrt() / (ref ** 2).sum(dim=(0, 1)).sqrt()
    print((B, Ci, Co, X, Y, Z), 'rel', e, 'db', float((db.double().cpu() - dyr.sum(dim=(0, 2, 3, 4))).norm() / dyr.sum(dim=(0,2,3,4)).norm()))
    if e > 1e-4: print(per_tap)
run(1, 48, 48, 4, 7, 160)
run(2, 48, 48, 24, 20, 28)
run(2, 96, 48, 6, 10, 10)
run(1, 192, 96, 3, 40, 40)
run(1, 768, 384, 5, 5, 5)
run(1, 48, 96, 9, 3, 21)
