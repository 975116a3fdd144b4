import sys, torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N
from nerf_mae_b200._lib import call
def run(B, Ci, Co, X, Y, Z, acc):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, X, Y, Z, Ci, generator=g).cuda()
    w = (torch.randn(Co, Ci, 3, 3, 3, generator=g) / (27 * Ci) ** 0.5).cuda()
    wws = torch.empty(27 * Ci * Co, device='cuda')
    y = torch.randn(B, X, Y, Z, Co, generator=g).cuda()
    y0 = y.clone()
    if acc:  # dgrad accumulate: x plays dout (Cout=Ci channels), result has Co channels -> weight shape (Cout=Ci, Cin=Co)
        w2 = (torch.randn(Ci, Co, 3, 3, 3, generator=g) / (27 * Ci) ** 0.5).cuda()
        call("nmae_conv3x3x3_dgrad", x, w2, B, X, Y, Z, Co, Ci, wws, y, 1, device=x.device)
        ref = torch.nn.functional.conv_transpose3d(x.permute(0, 4, 1, 2, 3).double(), w2.double(), padding=1).permute(0, 2, 3, 4, 1) + y0.double()
    else:
        call("nmae_conv3x3x3_fwd", x, w, None, B, X, Y, Z, Ci, Co, wws, y, device=x.device)
        ref = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), padding=1).permute(0, 2, 3, 4, 1)
    err = (y.double() - ref)
    e = float(err.norm() / ref.norm())
    # locate the bad rows
    bad = (err.abs().amax(dim=-1) > 1e-3 * float(ref.abs().max())).nonzero()
    print((B, Ci, Co, X, Y, Z, acc), 'rel', e, 'bad voxels', bad.shape[0], bad[:6].tolist(), bad[-3:].tolist())
run(1, 48, 48, 4, 7, 160, 0)
run(2, 48, 48, 24, 20, 28, 0)
run(2, 48, 48, 24, 20, 28, 1)
run(1, 48, 48, 4, 7, 160, 1)
run(1, 48, 48, 40, 160, 160, 0)
