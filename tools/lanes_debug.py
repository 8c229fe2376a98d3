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
NL = int(os.environ.get('NL', 3))
ctxs = [Context(0) for _ in range(NL)]
for c in ctxs:
    c.set_graph(indptr, indices, w); c.set_nodes(tr, te); c.set_targets(t[tr]); c.set_sm_budget(148 // NL)
def grab(c, which):
    p, nbytes = c.debug_d6_buffer(which)
    out = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(p), ctypes.c_size_t(nbytes), 3)
    torch.cuda.synchronize()
    return out.cpu().numpy()
ctxs[0].gpr_logp_grad(1, th, None, True); ref = ctxs[0].debug_scalars()
route, parts = ctxs[0].last_route(); so_guess = None
names = {0:'q11',1:'q1t',2:'qtt',3:'s1t',4:'stt',5:'fro',6:'froo',7:'ld',8:'ldo',9:'trT',10:'ldT',11:'trTo',12:'ldTo',13:'b1u1',14:'b1ut',15:'btut',16:'u1u1',17:'u1ut',18:'utut'}
BUFS = (0, 1, 2, 3, 4, 5, 7, 8)
refbuf = {w_: grab(ctxs[0], w_) for w_ in BUFS}
print("parts", parts)
for it in range(40):
    for c in ctxs: c.gpr_logp_grad_begin(1, th, None, True)
    sc = []
    for c in ctxs:
        c.gpr_logp_grad_end(); sc.append(c.debug_scalars())
    found = False
    for lane in range(NL):
        relv = np.abs(sc[lane][:19] - ref[:19]) / np.maximum(np.abs(ref[:19]), 1e-300)
        if relv.max() > 1e-12:
            print(it, "lane", lane, "scalars off:", [(names[i], f"{relv[i]:.1e}") for i in np.nonzero(relv > 1e-12)[0]])
            found = True
            break
    if found:
        for w_ in BUFS:
            x = grab(ctxs[lane], w_); r = refbuf[w_]
            d = np.abs(x - r)
            idx = np.nonzero(d > 1e-12 * np.max(np.abs(r)))[0]
            print("  buffer", w_, "size", x.size, "entries off:", len(idx), "max", d.max(), "first/last idx", (idx[:5], idx[-5:]) if len(idx) else None)
            if len(idx) and w_ == 3:
                ldso = 512 if x.size % 512 == 0 else 384
                for ld in (384, 512, 640):
                    if x.size % ld == 0: print("   ld", ld, "rows", np.unique(idx // ld)[:20], "cols", np.unique(idx % ld)[:20])
        # nature of the wrong rows of Y (buffer 2): stale input (TRSM applied to the previous evaluation's Y instead of U)?
        s_ = parts[1]; lds = (s_ + 127) // 128 * 128
        Yb = grab(ctxs[lane], 2).reshape(-1, lds)[:, :s_]; Yr = refbuf[2].reshape(-1, lds)[:, :s_]
        LT = np.tril(refbuf[4][: lds * lds].reshape(lds, lds)[:s_, :s_])
        rows = np.nonzero(np.max(np.abs(Yb - Yr), axis=1) > 1e-12 * np.max(np.abs(Yr)))[0]
        print("  bad rows of Y:", rows[:8], "...", rows[-8:], "count", len(rows))
        if len(rows):
            import scipy.linalg as sl
            Y2 = sl.solve_triangular(LT, Yr[rows].T, lower=True).T          # (Y L^-T) applied to the reference Y itself
            print("  |bad - TRSM(Y_ref)| max:", np.max(np.abs(Yb[rows] - Y2)), " |bad - Y_ref| max:", np.max(np.abs(Yb[rows] - Yr[rows])), " |bad| max:", np.max(np.abs(Yb[rows])))
            # per column tile, which columns differ
            cols = np.nonzero(np.max(np.abs(Yb[rows] - Yr[rows]), axis=0) > 1e-12 * np.max(np.abs(Yr)))[0]
            print("  bad cols:", cols[:6], "...", cols[-6:], "count", len(cols))
        break
print("done")
