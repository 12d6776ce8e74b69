"""Micro-benchmark of the solve-path kernels on synthetic stencil matrices (GPU box)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from parelag_b200 import capi
from tests.util import laplace3d

def stencil27(n):
    t = sp.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1])
    A = sp.kron(sp.kron(t, t), t).tocsr()
    A.data[:] = -1.0
    A = A + sp.diags(np.full(n ** 3, 27.0))
    return A.tocsr()

def bench(ctx, name, A, reps=20):
    n = A.shape[0]; nnz = A.nnz
    dA = capi.Mat.from_scipy(ctx, A)
    x = capi.Vec(ctx, data=np.random.default_rng(0).standard_normal(n)); y = capi.Vec(ctx, n); b = capi.Vec(ctx, data=np.ones(n))
    out = {"name": name, "n": n, "nnz": nnz}
    def timeit(f, nbytes):
        for _ in range(3): f()
        ts = []
        for _ in range(reps):
            ctx.flush_l2(); ctx.sync(); ctx.timer_start(); f(); ts.append(ctx.timer_stop())
        t = float(np.median(ts))
        return {"ms": t, "GBs": nbytes / t / 1e6}
    b_spmv = 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n
    out["spmv"] = timeit(lambda: dA.spmv(x, y), b_spmv)
    sj = capi.Smoother(ctx, dA, type=1)
    out["l1jacobi"] = timeit(lambda: sj.apply(b, x, True), b_spmv + 8 * n * 5)
    t0 = time.time(); sg = capi.Smoother(ctx, dA, type=2, ordering=capi.GS_MULTICOLOR); out["gs_mc_setup_s"] = time.time() - t0
    order, starts = sg.order(); out["colors"] = len(starts) - 1
    out["l1gs_multicolor_sym"] = timeit(lambda: sg.apply(b, x, True), 2 * (12 * nnz + 4 * n) + 2 * 32 * n)
    t0 = time.time(); sn = capi.Smoother(ctx, dA, type=2, ordering=capi.GS_NATURAL); out["gs_nat_setup_s"] = time.time() - t0
    order, starts = sn.order(); out["levels"] = len(starts) - 1
    out["l1gs_natural_sym"] = timeit(lambda: sn.apply(b, x, True), 2 * (12 * nnz + 4 * n) + 2 * 32 * n)
    print(json.dumps(out)); sys.stdout.flush()

if __name__ == "__main__":
    ctx = capi.Ctx()
    bench(ctx, "lap7_128", laplace3d(128, 128, 128))
    bench(ctx, "lap7_208", laplace3d(208, 208, 208))
    bench(ctx, "st27_160", stencil27(160))
