"""Shared helpers for the tests: synthetic matrices."""
import numpy as np
import scipy.sparse as sp


def laplace3d(nx, ny, nz):
    def l1(n):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    I = sp.identity
    A = (sp.kron(sp.kron(l1(nx), I(ny)), I(nz)) + sp.kron(sp.kron(I(nx), l1(ny)), I(nz))
         + sp.kron(sp.kron(I(nx), I(ny)), l1(nz)))
    return A.tocsr()


def random_spd(n, density, seed):
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=density, random_state=rng, format="csr")
    A = B + B.T
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    return A.tocsr()


def aggregation_P(nx, ny, nz, seed=0):
    """piecewise 'random weight' interpolation from 2x2x2 aggregates (RAP test input)."""
    rng = np.random.default_rng(seed)
    cx, cy, cz = (nx + 1) // 2, (ny + 1) // 2, (nz + 1) // 2
    rows, cols, vals = [], [], []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                r = (i * ny + j) * nz + k
                c = ((i // 2) * cy + j // 2) * cz + k // 2
                rows.append(r); cols.append(c); vals.append(rng.uniform(0.5, 1.5))
                if i + 1 < nx:   # a second column so rows of P overlap
                    c2 = (((i + 1) // 2) * cy + j // 2) * cz + k // 2
                    if c2 != c:
                        rows.append(r); cols.append(c2); vals.append(rng.uniform(-0.5, 0.5))
    return sp.csr_matrix((vals, (rows, cols)), shape=(nx * ny * nz, cx * cy * cz))


def sorted_csr(A):
    A = A.tocsr().copy()
    A.sort_indices()
    return A
