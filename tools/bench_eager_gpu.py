"""Side measurement for the north-star context ("vs the reference's 1-GPU PyTorch-eager grids/sec"): the reference ALGORITHM as
plain PyTorch eager ops on the same GPU - the oracle port (oracle/nerf_mae_oracle.py, the functional restatement that is pinned
to the live reference) with its tensors on cuda:0, swin_s, 4 x 160^3 grids, full train step (fwd+bwd+clip+AdamW), fp32, torch
defaults (cuDNN convolutions may use TF32, matmuls do not).  Not part of bench.py, tests or the product path.
    python tools/bench_eager_gpu.py [batch=4] [steps=3]"""
import random
import sys
import torch
sys.path.insert(0, '.')
from oracle import nerf_mae_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = O.SWIN_CONFIGS["swin_s"]
sd = {k: v.cuda() for k, v in O.init_state_dict("swin_s", 160, seed=0).items()}
for k, v in sd.items():
    v.requires_grad_(v.dtype.is_floating_point and k != "pos_embed")
g = torch.Generator().manual_seed(0)
grids = [torch.rand(4, 160, 160, 160, generator=g).cuda() for _ in range(B)]
torch.set_default_device("cuda")      # the oracle creates its index / mask / padding tensors without a device argument
state = {}
random.seed(0)
for i in range(1 + steps):
    if i == 1:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    O.train_step(sd, state, grids, cfg["depths"], cfg["num_heads"], 160, 0.75, lr=1e-4)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("PyTorch-eager (oracle port on cuda:0), swin_s, %d x 160^3, fp32, allow_tf32(conv)=%s: %.1f ms/step = %.2f grids/s, peak mem %.1f GB" % (
    B, torch.backends.cudnn.allow_tf32, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 1e9))
