"""Generate tests/golden/*.npz by running the UNMODIFIED reference (under oracle/ref_shim.py) on the
canonical inputs of oracle/configs.py, and check the numpy restatement (oracle/gp_oracle.py) against it
while doing so.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # all cases (C2-full diffusion takes ~1-2 min)

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import contextlib
import io
import os
import sys
import time

import numpy as np
import pandas as pd

from oracle import configs as cfg
from oracle import gp_oracle as go
from oracle.ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def csr_dict(G):
    indptr, indices, weights, _ = go.graph_to_csr(G)
    return dict(indptr=indptr, indices=indices, weights=weights)


def regressor_case(name, G, train, test, t, kerneltype, theta, REG, store_blocks, fd=True, landscape=None):
    t0 = time.time()
    tv = pd.Series(t[train], index=train)
    model = quiet(REG, Graph=G, training_nodes=list(train), test_nodes=list(test), theta=list(theta),
                  relabel_nodes=True, kerneltype=kerneltype, training_values=False)
    model.set_training_values(tv.copy())
    quiet(model.predict)
    logp = float(model.logp())
    out = dict(csr_dict(G), train=np.array(train), test=np.array(test), t_train=t[train], theta=np.array(theta),
               kerneltype=kerneltype, logp=logp, fstar=np.asarray(model.fstar), s=np.asarray(model.s),
               alpha=np.asarray(model.alpha), kstarstar_diag=np.asarray(model.kstarstar_diag))
    if store_blocks:
        out.update(k=np.asarray(model.k), kstar=np.asarray(model.kstar))
    if fd:   # central finite differences of the reference's own logPosterior (gradient pin)
        g = np.zeros(len(theta))
        for j in range(len(theta)):
            h = 1e-5
            tp, tm = np.array(theta, dtype=float), np.array(theta, dtype=float)
            tp[j] += h
            tm[j] -= h
            g[j] = (model.logPosterior(tp, train, tv) - model.logPosterior(tm, train, tv)) / (2 * h)
        out["neg_logp_grad_fd"] = g
    if landscape is not None:
        axidx, ax1, ax2 = landscape
        out.update(land_axidx=np.array(axidx), land_ax1=ax1, land_ax2=ax2,
                   land_lml=quiet(model.lml_landscape, list(theta), axidx, ax1, ax2))
    # ---- check the restatement against the reference
    orc = go.OracleGP(G, kerneltype=kerneltype)
    pr = orc.predict_regressor(theta, train, test, t[train])
    errs = dict(k=relerr(pr["k"], model.k), kstar=relerr(pr["kstar"], model.kstar),
                fstar=relerr(pr["fstar"], model.fstar), s=relerr(pr["s"], model.s),
                logp=abs(-orc.neg_logp(theta, train, t[train]) - logp) / abs(logp))
    if fd:
        _, ng = orc.neg_logp_grad(theta, train, t[train])
        out["neg_logp_grad_analytic_oracle"] = ng
        # p-step: theta[1] only enters through p = int(exp(theta[1])) (A-4) -> the FD probe straddles the
        # integer step and is meaningless there; the analytic derivative is 0.  Pin components 0 and -1.
        sel = [0, len(theta) - 1] if kerneltype == "pstep_walk" else list(range(len(theta)))
        out["grad_fd_valid"] = np.array(sel)
        errs["grad_vs_fd"] = relerr(ng[sel], out["neg_logp_grad_fd"][sel])
    if landscape is not None:
        errs["land"] = relerr(orc.landscape(theta, axidx, ax1, ax2, train, t[train]), out["land_lml"])
    bad = {k: v for k, v in errs.items() if not (v < (1e-5 if k == "grad_vs_fd" else 1e-10))}
    print(f"[{name}] {time.time()-t0:6.1f}s  logp={logp:.9f}  oracle-vs-ref {errs}")
    assert not bad, bad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def classifier_case(name, G, train, test, y, kerneltype, theta, CLS):
    t0 = time.time()
    yv = pd.Series(y[train].astype(np.int64), index=train)
    model = quiet(CLS, Graph=G, training_nodes=list(train), test_nodes=list(test), theta=list(theta),
                  relabel_nodes=True, kerneltype=kerneltype, training_values=yv)
    # count Newton iterations by wrapping np.linalg.cholesky calls inside NRiteration
    f, logq, a = model.NRiteration(train, yv, theta)
    fstar, V, probs = quiet(model.predict)
    labels = np.asarray(model.df["predicted_class"][test], dtype=np.float64)
    orc = go.OracleGP(G, kerneltype=kerneltype)
    pr = orc.predict_classifier(theta, train, test, y[train])
    out = dict(csr_dict(G), train=np.array(train), test=np.array(test), y_train=y[train], theta=np.array(theta),
               kerneltype=kerneltype, f=np.asarray(f).ravel(), a=np.asarray(a).ravel(), logq=float(np.squeeze(logq)),
               iters=pr["iters"], fstar=np.asarray(fstar).ravel(), V=np.asarray(V).ravel(), probs=np.asarray(probs),
               labels=labels, logp=float(np.squeeze(model.logp())))
    errs = dict(f=relerr(pr["f"].ravel(), out["f"]), a=relerr(pr["a"].ravel(), out["a"]),
                logq=abs(pr["logq"] - out["logq"]) / abs(out["logq"]), fstar=relerr(pr["fstar"], out["fstar"]),
                V=relerr(pr["V"], out["V"]), probs=relerr(pr["probs"], out["probs"]),
                labels=float(np.sum(pr["labels"] != labels)))
    print(f"[{name}] {time.time()-t0:6.1f}s  logq={out['logq']:.9f} iters={out['iters']} "
          f"labels(+1)={int((labels > 0).sum())}/{len(labels)}  oracle-vs-ref {errs}")
    bad = {k: v for k, v in errs.items() if not (v < 1e-9)}
    assert not bad, bad
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    quick = "--quick" in argv
    _, REG, CLS = load_reference()
    L = np.log
    thetas = {"diffusion": L([0.6, 0.01]), "regularized_laplacian": L([0.6, 0.01]), "pstep_walk": L([2.3, 4, 0.01])}
    short = {"diffusion": "diff", "regularized_laplacian": "reglap", "pstep_walk": "pstep"}
    # --- regressor, 15x15 grid (graph_test.py demo size): blocks + FD gradient + small landscape
    G, tr, te, t = cfg.grid_case(15, 15, 100, 100)
    for kt, th in thetas.items():
        land = ([0, len(th) - 1], np.linspace(L(0.05), L(0.95), 4) if kt != "pstep_walk" else np.linspace(L(2.05), L(3.0), 4),
                np.linspace(L(0.001), L(1.0), 3))
        regressor_case(f"reg_grid15_{short[kt]}", G, tr, te, t, kt, th, REG, store_blocks=True, landscape=land)
    # --- regressor, C1 (30x30 grid, 90/810)
    G, tr, te, t = cfg.grid_case(30, 30, 90, 810)
    for kt, th in thetas.items():
        regressor_case(f"reg_c1_grid30_{short[kt]}", G, tr, te, t, kt, th, REG, store_blocks=False, fd=(kt != "diffusion"))
    # --- regressor, reduced C3/C4 (RGG 1024, 512/512)
    G, tr, te, t = cfg.rgg_case(1024, 512, 512)
    for kt, th in thetas.items():
        regressor_case(f"reg_rgg1024_{short[kt]}", G, tr, te, t, kt, th, REG, store_blocks=False, fd=(kt != "diffusion"))
    # --- classifier, reduced C2 (BA 400) and the BA hub structure with all three kernels
    cth = {"diffusion": L([0.6, 0.1]), "regularized_laplacian": L([0.6, 0.1]), "pstep_walk": L([2.1, 2, 0.1])}
    G, tr, te, y = cfg.ba_case(400, 200, 200)
    for kt, th in cth.items():
        classifier_case(f"cls_ba400_{short[kt]}", G, tr, te, y, kt, th, CLS)
    G, tr, te, y = cfg.grid_case(15, 15, 125, 100, freq=0.6)
    y = np.where(y > 0, 1, -1)
    for kt, th in cth.items():
        classifier_case(f"cls_grid15_{short[kt]}", G, tr, te, y, kt, th if kt != "pstep_walk" else L([2.3, 4, 0.01]), CLS)
    if not quick:
        # --- C2 full size (BA 2000, 1000/1000)
        G, tr, te, y = cfg.ba_case(2000, 1000, 1000)
        for kt in ("pstep_walk", "regularized_laplacian", "diffusion"):
            classifier_case(f"cls_c2_ba2000_{short[kt]}", G, tr, te, y, kt, cth[kt], CLS)
        # --- reduced C3 at N=2048, reg-Laplacian (dense inverse: fast in the reference)
        G, tr, te, t = cfg.rgg_case(2048, 1024, 1024)
        regressor_case("reg_rgg2048_reglap", G, tr, te, t, "regularized_laplacian", thetas["regularized_laplacian"], REG,
                       store_blocks=False)
    # --- the notebook KAT (report_notebook.ipynb cell 49)
    K, yk = go.notebook_kat_inputs()
    f, logq, a, it = go.gpc_newton(K, yk, logq_variant="notebook")
    print("notebook KAT: -logq =", -logq, "iters", it, "expected", go.NOTEBOOK_KAT_VALUE)
    assert abs(-logq - go.NOTEBOOK_KAT_VALUE) < 5e-9
    np.savez_compressed(os.path.join(OUT, "notebook_kat.npz"), K=K, y=yk.ravel(), neg_logq=go.NOTEBOOK_KAT_VALUE,
                        iters=it, f=f.ravel(), a=a.ravel())


if __name__ == "__main__":
    main(sys.argv[1:])
