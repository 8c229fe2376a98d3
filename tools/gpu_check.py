"""Bring-up checks on a real B200 (run stage by stage under gpurun): python tools/gpu_check.py <stage> ...
Stages: gemm potrf inverse laplacian kernels gpr gpc perf.  Prints max relative errors against torch fp64 /
the numpy oracle.  Development tool; the judged tests are tests/ (-m gpu)."""
import glob
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpnet_b200._lib import Context, KERNEL_IDS  # noqa: E402
from oracle import configs as cfg  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

dev = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def dmat(rows, cols, fill=None):
    ld = (cols + 15) // 16 * 16
    buf = torch.zeros(rows, ld, dtype=torch.float64, device=dev)
    if fill is not None:
        buf[:, :cols] = fill
    return buf, ld


def ev(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def stage_gemm(ctx):
    torch.manual_seed(0)
    stream = torch.cuda.current_stream().cuda_stream
    for (M, N, K) in [(128, 128, 128), (64, 64, 16), (300, 200, 100), (1000, 900, 333), (1024, 1024, 1024), (2048, 2048, 512),
                      (4096, 4096, 256)]:
        for ta in (0, 1):
            for tb in (0, 1):
                A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device=dev)
                B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device=dev)
                Ab, lda = dmat(A.shape[0], A.shape[1], A)
                Bb, ldb = dmat(B.shape[0], B.shape[1], B)
                C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
                Cb, ldc = dmat(M, N, C0)
                ctx.gemm(ta, tb, M, N, K, 0.75, Ab.data_ptr(), lda, Bb.data_ptr(), ldb, -0.5, Cb.data_ptr(), ldc)
                ctx.sync()
                ref = 0.75 * ((A.T if ta else A) @ (B.T if tb else B)) - 0.5 * C0
                err = (Cb[:, :N] - ref).abs().max().item() / ref.abs().max().item()
                pad = Cb[:, N:].abs().max().item() if ldc > N else 0.0
                print(f"gemm M{M} N{N} K{K} ta{ta} tb{tb}: relerr {err:.2e} pad-touched {pad:.1e}", flush=True)
    del stream


def stage_perf(ctx):
    for n in (4096, 8192):
        A, ld = dmat(n, n, torch.randn(n, n, dtype=torch.float64, device=dev))
        B, _ = dmat(n, n, torch.randn(n, n, dtype=torch.float64, device=dev))
        Cb, _ = dmat(n, n)
        for ta, tb in ((0, 1), (0, 0), (1, 0)):
            ms = ev(lambda: (ctx.gemm(ta, tb, n, n, n, 1.0, A.data_ptr(), ld, B.data_ptr(), ld, 0.0, Cb.data_ptr(), ld), ctx.sync()))
            print(f"gemm {n}^3 ta{ta} tb{tb}: {ms:.3f} ms  {2*n**3/ms*1e-9:.2f} TFLOP/s", flush=True)
        ms = ev(lambda: torch.matmul(A[:, :n], B[:, :n]))
        print(f"cublas {n}^3: {ms:.3f} ms {2*n**3/ms*1e-9:.2f} TFLOP/s", flush=True)
        X = torch.randn(n, n, dtype=torch.float64, device=dev)
        S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        Sb, ld = dmat(n, n)
        Ob, _ = dmat(n, n)
        Wb, _ = dmat(n, n)

        def potrf():
            Sb[:, :n] = S
            ctx.potrf(n, Sb.data_ptr(), ld)

        def copy_only():
            Sb[:, :n] = S
            ctx.sync()

        def inv():
            Sb[:, :n] = S
            ctx.spd_inverse(n, Sb.data_ptr(), ld, Ob.data_ptr(), ld, Wb.data_ptr())

        t_copy = ev(copy_only)
        t_p = ev(potrf) - t_copy
        t_i = ev(inv) - t_copy
        print(f"potrf {n}: {t_p:.3f} ms {n**3/3/t_p*1e-9:.2f} TFLOP/s | spd_inverse {n}: {t_i:.3f} ms {n**3/t_i*1e-9:.2f} TFLOP/s (N^3 flops)",
              flush=True)
        ms = ev(lambda: torch.linalg.cholesky(S))
        print(f"cusolver potrf {n}: {ms:.3f} ms", flush=True)


def stage_leaf(ctx):
    n = 128
    X = torch.randn(n, n, dtype=torch.float64, device=dev)
    S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
    bufs = [dmat(n, n, S)[0] for _ in range(50)]

    def run():
        for b in bufs:
            ctx.lib.gpn_potrf_f64(ctx.h, n, b.data_ptr(), 128, None)   # async (no info read-back)
    import ctypes
    ctx.lib.gpn_potrf_f64.argtypes[4] = ctypes.c_void_p
    ms = ev(lambda: (run(), ctx.sync()))
    print(f"leaf potf2+inverse 128x128: {ms/50*1e3:.1f} us per launch (incl. zero_upper + launch overhead)", flush=True)
    for copies in (1, 148):
        stack = S.repeat(copies, 1).contiguous()
        prof = torch.zeros(32, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        ctx.lib.gpn_debug_leaf_bench(ctx.h, stack.data_ptr(), copies, prof.data_ptr())
        ctx.sync()
        p = prof.cpu().numpy()
        stamps = [int(x) for x in p[:12] if x != 0]
        print(f"leaf phase cycles, grid {copies} (start, D0 + lockstep inverse, then [panel product + inverse row | next diagonal update + D(kb+1)] x 3, last inverse row, last write-back):")
        print("   phases:", [b - a for a, b in zip(stamps[:-1], stamps[1:])], " total:", stamps[-1] - stamps[0], flush=True)
        if p[12] != 0:       # v2 diagnostics of phase (c), cycles after the barrier that starts it
            for kb in range(3):
                t0 = int(p[12 + kb])
                print(f"   kb={kb}: factor warp done +{int(p[16+kb])-t0}, lockstep inverse done +{int(p[20+kb])-t0}, trailing update + T done "
                      f"+{int(p[24+kb])-t0}, write-back (warp 4) done +{int(p[28+kb])-t0}; next-diagonal update before it: {t0 - stamps[2+2*kb]}", flush=True)
    for nn in (256, 512, 1024, 2048):
        X = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        S = X @ X.T / nn + torch.eye(nn, dtype=torch.float64, device=dev)
        Sb, ld = dmat(nn, nn, S)
        ms = ev(lambda: (ctx.potrf(nn, Sb.data_ptr(), ld)))
        print(f"potrf {nn}: {ms*1e3:.1f} us  ({nn//128} tile columns -> {ms*1e3/(nn//128):.1f} us per column)", flush=True)


def stage_chain(ctx):
    """Per-task %globaltimer trace of the dataflow Cholesky: where the per-column dependency chain spends its time."""
    for n in (2048, 8192):
        T = n // 128
        X = torch.randn(n, n, dtype=torch.float64, device=dev)
        S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        Sb, ld = dmat(n, n, S)
        ctx.potrf(n, Sb.data_ptr(), ld)
        Sb[:, :n] = S
        ntasks = T * T                                  # trace slot of tile (i, j): j * T + i
        tr = torch.zeros(ntasks * 8, dtype=torch.int64, device=dev)
        ctx.lib.gpn_debug_df_trace(ctx.h, tr.data_ptr())
        ctx.potrf(n, Sb.data_ptr(), ld)
        ctx.lib.gpn_debug_df_trace(ctx.h, None)
        t = tr.cpu().numpy().reshape(ntasks, 8).astype(np.float64) * 1e-3      # us
        q0 = lambda j: j * T + j
        rows = []
        for j in range(T - 1):
            dg, ts, dn = t[q0(j)], t[q0(j) + 1], t[q0(j + 1)]     # diag(j), trsm(j+1,j), diag(j+1)
            rows.append([dg[4] - dg[2],          # leaf (S staged -> flag)
                         dg[2] - dg[1],          # diag epilogue (A read, S staging)
                         ts[3] - dg[4],          # flag(j,j) -> first W stage landed in the TRSM CTA
                         ts[5] - ts[3],          # TRSM GEMM
                         ts[4] - ts[5],          # TRSM store + fence + flag
                         dn[1] - ts[4],          # flag(j+1,j) -> last k-block of diag(j+1) done
                         dn[4] - dg[4]])         # column period
        r = np.array(rows)
        names = ["leaf", "diag-epilogue", "flag->W landed", "trsm gemm", "trsm store+flag", "flag->lastk done", "column period"]
        live = t[:, 4] > 0
        print(f"n={n}: total {(t[live, 4].max() - t[live, 0].min()):.0f} us;  chain segments (us, median over columns; [first 5] .. [last 5]):")
        for k, nm in enumerate(names):
            print(f"   {nm:18s} median {np.median(r[:, k]):7.1f}   head {np.round(r[:5, k], 1)}  tail {np.round(r[-5:, k], 1)}")
        sys.stdout.flush()


def stage_dfbreak(ctx):
    """Where the CTA-time of the dataflow Cholesky goes (per-task %globaltimer trace, N = 8192, 148 CTAs): main loops,
    epilogues (TRSM / leaf), waiting between tasks, idle after a CTA's last task."""
    n = 8192
    T = n // 128
    X = torch.randn(n, n, dtype=torch.float64, device=dev)
    S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
    Sb, ld = dmat(n, n, S)
    ctx.potrf(n, Sb.data_ptr(), ld)
    Sb[:, :n] = S
    tr = torch.zeros(T * T * 8, dtype=torch.int64, device=dev)
    ctx.lib.gpn_debug_df_trace(ctx.h, tr.data_ptr())
    ctx.potrf(n, Sb.data_ptr(), ld)
    ctx.lib.gpn_debug_df_trace(ctx.h, None)
    full = tr.cpu().numpy().reshape(T * T, 8).astype(np.float64) * 1e-3      # us; slot of tile (i, j): j * T + i
    order = [j * T + i for j in range(T) for i in range(j, T)]                # the kernel's task order (dealt round-robin)
    t = full[order]
    ntasks = len(order)
    G = 148
    t0, t1 = t[:, 0].min(), t[:, 4].max()
    tot = (t1 - t0) * G
    main = epi = gap = tail = 0.0
    jof = np.array([s // T for s in order])
    ideal = 0.0
    for cta in range(G):
        prev = t0
        for q in range(cta, ntasks, G):
            gap += t[q, 0] - prev
            main += t[q, 1] - t[q, 0]
            epi += t[q, 4] - t[q, 1]
            prev = t[q, 4]
            ideal += jof[q] * 2 * 128 ** 3 / (37.0e12 / 148) * 1e6      # off-diagonal k-block at the per-SM DMMA ceiling
        tail += t1 - prev
    print(f"potrf {n}: {t1 - t0:.0f} us on {G} CTAs; CTA-time: main loops {100*main/tot:.1f} % (at the DMMA ceiling they would take "
          f"{100*ideal/tot:.1f} %), epilogues {100*epi/tot:.1f} %, start-of-task gaps {100*gap/tot:.1f} %, idle after last task {100*tail/tot:.1f} %")
    # the same per phase of the factorisation: thirds of the elapsed time
    for lo, hi in ((0, 1 / 3), (1 / 3, 2 / 3), (2 / 3, 1)):
        a, b = t0 + lo * (t1 - t0), t0 + hi * (t1 - t0)
        busy = 0.0
        for q in range(ntasks):
            busy += max(0.0, min(t[q, 1], b) - max(t[q, 0], a))
        print(f"   elapsed {lo:.2f}-{hi:.2f}: main-loop occupancy {100*busy/((b-a)*G):.1f} % of CTA-time")
    sys.stdout.flush()


def stage_potrf(ctx):
    torch.manual_seed(1)
    for n in (5, 100, 128, 129, 256, 900, 1024, 2048, 3000):
        X = torch.randn(n, n, dtype=torch.float64, device=dev)
        S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        Sb, ld = dmat(n, n, S)
        info = ctx.potrf(n, Sb.data_ptr(), ld)
        L = torch.linalg.cholesky(S)
        err = (Sb[:, :n] - L).abs().max().item() / L.abs().max().item()
        print(f"potrf n{n}: info {info} relerr {err:.2e}", flush=True)
    # not PD
    n = 300
    S = torch.eye(n, dtype=torch.float64, device=dev)
    S[200, 200] = -1.0
    Sb, ld = dmat(n, n, S)
    print("potrf non-PD info (expect 201):", ctx.potrf(n, Sb.data_ptr(), ld), flush=True)


def stage_inverse(ctx):
    torch.manual_seed(2)
    for n in (7, 128, 300, 900, 1024, 2048, 2500):
        X = torch.randn(n, n, dtype=torch.float64, device=dev)
        S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        Sb, ld = dmat(n, n, S)
        Ob, _ = dmat(n, n)
        Wb, _ = dmat(n, n)
        info = ctx.spd_inverse(n, Sb.data_ptr(), ld, Ob.data_ptr(), ld, Wb.data_ptr())
        ref = torch.linalg.inv(S)
        err = (Ob[:, :n] - ref).abs().max().item() / ref.abs().max().item()
        resid = (Ob[:, :n] @ S - torch.eye(n, dtype=torch.float64, device=dev)).abs().max().item()
        print(f"spd_inverse n{n}: info {info} relerr {err:.2e} resid {resid:.2e}", flush=True)


def stage_laplacian(ctx):
    G, tr, te, t = cfg.grid_case(15, 15, 100, 100)
    G.add_edge(3, 3)            # a self loop
    G.add_node(225)             # an isolated node (d = 0 -> dinv = 0)
    G[0][1]["weight"] = 2.5     # a weighted edge
    indptr, indices, w, _ = go.graph_to_csr(G)
    N = len(indptr) - 1
    import networkx as nx
    Lnx = nx.normalized_laplacian_matrix(G).toarray()
    Lor = go.normalized_laplacian(indptr, indices, w, N)
    print("oracle vs networkx laplacian:", np.max(np.abs(Lor - Lnx)))
    ctx.set_graph(indptr, indices, w)
    out, ld = dmat(N, N)
    ctx.laplacian_dense(1.0, 0.0, out.data_ptr(), ld)
    ctx.sync()
    print("gpu laplacian vs networkx: maxabs", np.max(np.abs(out[:, :N].cpu().numpy() - Lnx)), flush=True)
    ctx.laplacian_dense(-0.6, 0.0, out.data_ptr(), ld)
    ctx.sync()
    print("gpu -0.6L bit-exact:", np.array_equal(out[:, :N].cpu().numpy(), -0.6 * Lnx), flush=True)
    ctx.laplacian_dense(0.6, 1.0, out.data_ptr(), ld)
    ctx.sync()
    print("gpu I+0.6L bit-exact:", np.array_equal(out[:, :N].cpu().numpy(), np.eye(N) + 0.6 * Lnx), flush=True)
    rows = torch.tensor(tr, dtype=torch.int32, device=dev)
    cols = torch.tensor(te, dtype=torch.int32, device=dev)
    blk, ldb = dmat(len(tr), len(te))
    ctx.gather_block_add(out.data_ptr(), ld, rows.data_ptr(), len(tr), cols.data_ptr(), len(te), 0.01, blk.data_ptr(), ldb)
    ctx.sync()
    ref = (np.eye(N) + 0.6 * Lnx)[np.ix_(tr, te)] + 0.01
    print("gather bit-exact:", np.array_equal(blk[:, :len(te)].cpu().numpy(), ref), flush=True)


def stage_kernels(ctx):
    L = np.log
    for name, (G, tr, te, t) in {"grid15": cfg.grid_case(15, 15, 100, 100), "grid30": cfg.grid_case(30, 30, 90, 810),
                                 "ba400": cfg.ba_case(400, 200, 200)}.items():
        indptr, indices, w, _ = go.graph_to_csr(G)
        N = len(indptr) - 1
        ctx.set_graph(indptr, indices, w)
        Lnorm = go.normalized_laplacian(indptr, indices, w, N)
        for kt, th in (("diffusion", L([0.6, 0.01])), ("diffusion", L([0.05, 0.01])), ("diffusion", L([0.99, 0.01])),
                       ("regularized_laplacian", L([0.6, 0.01])), ("pstep_walk", L([2.3, 4, 0.01])), ("pstep_walk", L([2.1, 2, 0.1])),
                       ("pstep_walk", L([2.5, 7.5, 0.1])), ("pstep_walk", L([2.5, 1.5, 0.1]))):
            for wg in (0, 1):
                st = ctx.kernel_build(KERNEL_IDS[kt], th, wg)
                K = ctx.kernel_block(range(N), range(N), 0.0)
                ref = go.full_kernel(Lnorm, kt, th, expm_mode="dense")
                extra = f" pade(m,s)={ctx.expm_info()}" if kt == "diffusion" else ""
                print(f"{name} {kt} theta={np.exp(th)} grad{wg}: st {st} relerr {rel(K, ref):.2e} sym {np.max(np.abs(K-K.T)):.1e}{extra}",
                      flush=True)


def stage_gpr(ctx):
    for path in sorted(glob.glob(os.path.join(GOLD, "reg_*.npz"))):
        g = np.load(path, allow_pickle=True)
        kt = str(g["kerneltype"])
        ctx.set_graph(g["indptr"], g["indices"], g["weights"])
        ctx.set_nodes(g["train"], g["test"])
        st, out, info = ctx.gpr_logp_grad(KERNEL_IDS[kt], g["theta"], g["t_train"], True)
        msg = f"{os.path.basename(path)}: st {st} info {info} logp relerr {abs(-out[0]-g['logp'])/abs(g['logp']):.2e}"
        if "neg_logp_grad_fd" in g:
            sel = g["grad_fd_valid"]
            msg += f" grad-vs-refFD {rel(out[1:][sel], g['neg_logp_grad_fd'][sel]):.2e} grad-vs-oracle {rel(out[1:], g['neg_logp_grad_analytic_oracle']):.2e}"
        st, out0, info = ctx.gpr_logp_grad(KERNEL_IDS[kt], g["theta"], g["t_train"], False)
        msg += f" | nograd logp relerr {abs(-out0[0]-g['logp'])/abs(g['logp']):.2e}"
        st, pr = ctx.gpr_predict(KERNEL_IDS[kt], g["theta"], g["t_train"])
        msg += f" | predict fstar {rel(pr['fstar'], g['fstar']):.2e} s {rel(pr['s'], g['s']):.2e} alpha {rel(pr['alpha'], g['alpha']):.2e} infos {pr['info_k']},{pr['info_kss']}"
        if "k" in g:
            msg += f" k {rel(ctx.gpr_fetch('k'), g['k']):.2e} kstar {rel(ctx.gpr_fetch('kstar'), g['kstar']):.2e}"
        print(msg, flush=True)


def stage_gpc(ctx):
    g = np.load(os.path.join(GOLD, "notebook_kat.npz"))
    n = len(g["y"])
    Kb, ld = dmat(n, n, torch.tensor(g["K"], device=dev))
    st, r = ctx.gpc_newton(0, [0.0], g["y"], variant=1, K_ptr=Kb.data_ptr(), ldk=ld)
    print(f"notebook KAT: -logq {-r['logq']:.8f} (expect {float(g['neg_logq']):.8f}) iters {r['iters']} (expect {int(g['iters'])})"
          f" f relerr {rel(r['f'], g['f']):.2e}", flush=True)
    for path in sorted(glob.glob(os.path.join(GOLD, "cls_*.npz"))):
        g = np.load(path, allow_pickle=True)
        kt = str(g["kerneltype"])
        ctx.set_graph(g["indptr"], g["indices"], g["weights"])
        ctx.set_nodes(g["train"], g["test"])
        t0 = time.time()
        st, r = ctx.gpc_predict(KERNEL_IDS[kt], g["theta"], g["y_train"])
        dt = time.time() - t0
        probs = np.vstack((1 - r["pi_star"], r["pi_star"])).T
        labels = go.predicted_class(probs)
        print(f"{os.path.basename(path)}: st {st} info {r['info']} iters {r['iters']}/{int(g['iters'])} logq {abs(r['logq']-g['logq'])/abs(g['logq']):.2e}"
              f" f {rel(r['f'], g['f']):.2e} a {rel(r['a'], g['a']):.2e} fstar {rel(r['fstar'], g['fstar']):.2e} V {rel(r['V'], g['V']):.2e}"
              f" probs {rel(probs, g['probs']):.2e} label-mismatch {int(np.sum(labels != g['labels']))} ({dt:.2f}s)", flush=True)


def stage_augbreak(ctx):
    """CTA-time breakdown of the augmented dataflow kernels of the coupled regularized-Laplacian route (N = 8192, n = 4096):
    per-task %globaltimer trace of the LAST dataflow launch ([S; X^T; I]) of one logp+grad evaluation.  Tasks are attributed
    to CTAs by their SM-independent time stamps: the kernel runs G CTAs, each a chain of non-overlapping tasks."""
    import oracle.configs as cfg
    N, n = 8192, 4096
    indptr, indices, w, pos, t = cfg.rgg_csr_fast(N)
    ctx.set_graph(indptr, indices, w)
    tr, te = cfg.split(N, n, N - n)
    ctx.set_nodes(tr, te)
    th = np.log([0.6, 0.01])
    ctx.gpr_logp_grad(1, th, t[tr], True)
    T, Trows, id0 = n // 128, 3 * (n // 128), 2 * (n // 128)
    trc = torch.zeros(T * Trows * 8, dtype=torch.int64, device=dev)
    ctx.lib.gpn_debug_df_trace(ctx.h, trc.data_ptr())
    ctx.gpr_logp_grad(1, th, t[tr], True)
    ctx.lib.gpn_debug_df_trace(ctx.h, None)
    tt = trc.cpu().numpy().reshape(T, Trows, 8).astype(np.float64) * 1e-3          # [j][i]
    tasks = [(i, j) for j in range(T) for i in range(j, Trows) if not (i >= id0 and j < i - id0)]
    G = 148
    st = np.array([tt[j, i, 0] for (i, j) in tasks]); ml = np.array([tt[j, i, 1] for (i, j) in tasks]); en = np.array([tt[j, i, 4] for (i, j) in tasks])
    t0, t1 = st.min(), en.max()
    tot = (t1 - t0) * G
    main, epi = float(np.sum(ml - st)), float(np.sum(en - ml))
    print(f"[S; X^T; I] n={n}: {t1 - t0:.0f} us; CTA-time: main loops (incl. waiting for operands) {100*main/tot:.1f} %, epilogues {100*epi/tot:.1f} %, "
          f"between / after tasks {100*(tot-main-epi)/tot:.1f} %")
    kb = sum((j - (i - id0 if i >= id0 else 0)) for (i, j) in tasks)
    print(f"   k-blocks {kb}: at 17 us each on {G} CTAs: {kb*17/G:.0f} us")
    for a in range(8):
        lo, hi = t0 + a / 8 * (t1 - t0), t0 + (a + 1) / 8 * (t1 - t0)
        busy = float(np.sum(np.clip(np.minimum(ml, hi) - np.maximum(st, lo), 0, None)))
        dg = [j for (i, j), e in zip(tasks, en) if i == j and lo <= e < hi]
        print(f"   elapsed {a}/8: main-loop occupancy {100*busy/((hi-lo)*G):5.1f} %   diagonal tiles finished: {min(dg) if dg else '-'}..{max(dg) if dg else '-'}")
    # the tail: tasks that end in the last eighth of the run
    lo = t0 + 7 / 8 * (t1 - t0)
    late = [(en[q] - t0, st[q] - t0, ml[q] - st[q], en[q] - ml[q], tasks[q]) for q in range(len(tasks)) if en[q] >= lo]
    late.sort()
    print(f"   {len(late)} tasks end in the last eighth; the last 12 (end, start, main loop, epilogue in us; tile (i, j); block):")
    for e, s0, m, ep, (i, j) in late[-12:]:
        print(f"      end {e:7.0f}  start {s0:7.0f}  main {m:6.0f}  epilogue {ep:5.0f}   ({i:3d},{j:3d})  {'square' if i < T else ('rhs' if i < id0 else 'identity')}")
    bycol = {}
    for e, s0, m, ep, (i, j) in late:
        bycol[j] = bycol.get(j, 0) + 1
    print("   tasks ending in the last eighth by tile column:", dict(sorted(bycol.items())))
    sys.stdout.flush()


def stage_potrftime(ctx):
    """potrf timing only (for sweeps over GPNET_B200_POTRF_GRID and the like)."""
    for n in (4096, 8192):
        X = torch.randn(n, n, dtype=torch.float64, device=dev)
        S = X @ X.T / n + torch.eye(n, dtype=torch.float64, device=dev)
        Sb, ld = dmat(n, n)

        def potrf():
            Sb[:, :n] = S
            ctx.potrf(n, Sb.data_ptr(), ld)

        def copy_only():
            Sb[:, :n] = S
            ctx.sync()

        t_p = ev(potrf, reps=5) - ev(copy_only, reps=5)
        print(f"potrf {n}: {t_p:.3f} ms {n**3/3/t_p*1e-9:.2f} TFLOP/s", end="   ")
    print(flush=True)


if __name__ == "__main__":
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)            # torch ops and library kernels share one (non-default) stream
    ctx = Context(0, ts.cuda_stream)
    for s in sys.argv[1:]:
        print(f"=== {s} ===", flush=True)
        globals()["stage_" + s](ctx)
