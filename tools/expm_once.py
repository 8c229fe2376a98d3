"""Two diffusion kernel builds (expm by Pade + squaring) at N = 8192 on the C3 graph; under ncu the second one is captured."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpnet_b200._lib import Context, KERNEL_IDS
from bench import c3_problem
N = int(os.environ.get("N", 8192))
G, csr, train, test, t = c3_problem(N)
ctx = Context(0)
ctx.set_graph(*csr)
for _ in range(2):
    ctx.kernel_build(KERNEL_IDS["diffusion"], np.log([0.6, 0.01]), False)
ctx.sync()
print("expm order / squarings:", ctx.expm_info(), "launches", ctx.launch_count())
