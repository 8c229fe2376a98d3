"""Profile target: one warm-up + one measured potrf / spd_inverse at size n (argv[1]), for ncu launch lists."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpnet_b200._lib import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "potrf"
dev = torch.device("cuda:0")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
ctx = Context(0, ts.cuda_stream)
X = torch.randn(n, n, dtype=torch.float64, device=dev)
S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
Sb = S.clone(); Ob = torch.empty_like(S); Wb = torch.empty_like(S)
for it in range(2):
    Sb.copy_(S)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if what == "potrf":
        info = ctx.potrf(n, Sb.data_ptr(), n)
    else:
        info = ctx.spd_inverse(n, Sb.data_ptr(), n, Ob.data_ptr(), n, Wb.data_ptr())
    e1.record(); torch.cuda.synchronize()
    print(what, n, "info", info, "ms", e0.elapsed_time(e1), "launches", ctx.launch_count())
