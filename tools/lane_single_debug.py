import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpnet_b200._lib import Context
from gpnet_b200.engine import graph_csr
from oracle import configs as cfg
N = 8192
G, tr, te, t = cfg.rgg_case(N, N // 2, N // 2)
indptr, indices, w, _, _ = graph_csr(G)
th = np.log([0.6, 0.01])
c = Context(0)
c.set_graph(indptr, indices, w); c.set_nodes(tr, te); c.set_targets(t[tr])
def grab(which):
    p, nbytes = c.debug_d6_buffer(which)
    out = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(p), ctypes.c_size_t(nbytes), 3)
    torch.cuda.synchronize()
    return out.cpu().numpy()
BUFS = (0, 1, 7, 8, 4, 5, 2, 3, 9, 10, 11)
prev = None
for it in range(6):
    st, out, info = c.gpr_logp_grad(1, th, None, True)
    cur = {b: grab(b) for b in BUFS}
    if prev is not None:
        msg = []
        for b in BUFS:
            d = np.abs(cur[b] - prev[b]); n = int(np.sum(d > 1e-13 * max(np.max(np.abs(prev[b])), 1e-300)))
            if n: msg.append((b, n, float(d.max()), int(np.argmax(d))))
        print(it, "buffers differing from previous evaluation (id, count, max, argmax):", msg)
    prev = cur
L = c.debug_d6_layout(); print(L)
s_, ND = L["s"], L["ND"]; lds = (s_ + 127) // 128 * 128; ldg = (ND + 127) // 128 * 128
GT = prev[0][: lds * ldg].reshape(lds, ldg)
for k in range(L["K"]):
    a, b = L["off"][k] + L["c0"][k], L["off"][k] + L["nk"][k]
    ref = GT[:s_, a:b] @ GT[:s_, a:b].T
    got = prev[7][k * lds * lds:(k + 1) * lds * lds].reshape(lds, lds)[:s_, :s_]
    d = np.abs(np.tril(got - ref))
    got2 = prev[9][k * lds * lds:(k + 1) * lds * lds].reshape(lds, lds)[:s_, :s_]
    GS = prev[10][: lds * ldg].reshape(lds, ldg)
    print("part", k, "Cb vs second copy:", np.abs(np.tril(got - got2)).max(), " GT snapshot vs final GT:", np.abs(GS[:s_, a:b] - GT[:s_, a:b]).max(),
          " second copy vs host:", np.abs(np.tril(got2 - ref)).max())
    got3 = prev[11][k * lds * lds:(k + 1) * lds * lds].reshape(lds, lds)[:s_, :s_]
    print("   plain-load copy vs host:", np.abs(np.tril(got3 - ref)).max())
    dd = np.abs(np.tril(got2 - ref)); bad = np.argwhere(dd > 1e-12)
    if len(bad):
        rows_, cols_ = np.unique(bad[:, 0]), np.unique(bad[:, 1])
        print("   wrong entries:", len(bad), "rows", rows_[:12], "..", rows_[-4:], "cols", cols_[:12], "..", cols_[-4:])
        tiles = sorted(set((int(r) // 64, int(c_) // 64) for r, c_ in bad)); print("   64-tiles:", tiles[:20])
        r0 = bad[0][0]; print("   row", r0, "bad cols:", bad[bad[:, 0] == r0][:, 1][:40], " rel err:", (dd[r0][dd[r0] > 1e-12] / np.abs(ref[r0][dd[r0] > 1e-12]))[:8])
    print("part", k, "Cb vs host GT GT^T (lower): max", d.max(), "at", np.unravel_index(np.argmax(d), d.shape), "scale", np.abs(ref).max())
for b in range(12):
    p, nb_ = c.debug_d6_buffer(b)
    if p: print("work", b, hex(p), "..", hex(p + nb_), nb_)
