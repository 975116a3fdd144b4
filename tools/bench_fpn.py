"""BASELINE config 5 on one GPU: swin_l encoder-only + FPN feature extraction, synthetic 160^3 x 4 grids, inference.
    python tools/bench_fpn.py [batch=8] [steps=5] [backbone=swin_l]
Prints grids/s (CUDA events, inputs resident in HBM) - a side measurement, the round's bench line is bench.py."""
import sys
import torch
sys.path.insert(0, '.')
import nerf_mae_b200 as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
backbone = sys.argv[3] if len(sys.argv) > 3 else "swin_l"
torch.manual_seed(0)
m = N.SwinTransformer_FPN_Pretrained_Skip(resolution=160, is_eval=True, backbone_type=backbone).cuda().eval()
m.fpn_neck.init_weights()
x = torch.rand(B, 4, 160, 160, 160, device="cuda")
with torch.no_grad():
    for _ in range(2):
        outs = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        outs = m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("config 5 (%s encoder + FPN, 160^3, batch %d, fp32, inference): %.1f ms/batch = %.1f grids/s; outputs %s" % (
    backbone, B, ms, B / ms * 1e3, [tuple(o.shape) for o in outs]))
