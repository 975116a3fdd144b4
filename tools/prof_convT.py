"""Times the decoder1 transposed convolution (kernel == stride == 4, 96 -> 48 channels, 40^3 -> 160^3, B=4) forward and backward in
isolation; used under ncu as well."""
import os
import sys

import torch

sys.path.insert(0, '.')
import nerf_mae_b200 as N  # noqa: F401
from nerf_mae_b200 import _lib

if os.environ.get("NMAE_USE_DBG_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libnmae_dbg.so")
if os.environ.get("NMAE_LIB_PATH"):     # same-box A/B against another build of the library
    _lib.LIB_PATH = os.environ["NMAE_LIB_PATH"]
call = _lib.call
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, X, Cin, Cout, k = 4, 40, 96, 48, 4
dev = torch.device("cuda")
x = torch.randn(B, X, X, X, Cin, device=dev)
w = torch.randn(Cin, Cout, k, k, k, device=dev) * 0.02
b = torch.zeros(Cout, device=dev)
out = torch.empty(B, X * k, X * k, X * k, Cout, device=dev)
dout = torch.randn_like(out)
dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
ws = torch.empty(w.numel(), device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for i in range(reps):
    ev[0].record()
    call("nmae_convT_k_eq_s_fwd", x, w, b, B, X, X, X, Cin, Cout, k, out, Cout, ws, device=dev)
    ev[1].record()
    call("nmae_convT_k_eq_s_bwd", dout, Cout, x, w, B, X, X, X, Cin, Cout, k, dx, dw, db, ws, device=dev)
    ev[2].record()
torch.cuda.synchronize()
print("convT k=s=4 96->48 40^3->160^3 B=4: fwd %.3f ms (output %.2f GB), bwd %.3f ms" % (
    ev[0].elapsed_time(ev[1]), out.numel() * 4 / 1e9, ev[1].elapsed_time(ev[2])))
